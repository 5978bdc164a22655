// lsw_chain.cuh -- the "chain" step that consumes the seed hits, on the device:
//   removeOverlappingPrimerAlignments   /root/reference/lamprey.py:859-899    (prune_overlaps_read)
//   extractAmpliconAroundTarget         /root/reference/lamprey.py:903-998  \  one (start, end, direction) per surviving
//   stage 1 of processLamplicon         /root/reference/lamprey.py:1566-1589 /  target hit of a read (chain_subreads_read)
// Both walk one read's hits in their final (start, search order) order; one thread owns one read.  The reference's quirks are
// kept: the leftward walk never looks at the read's first surviving hit (`range(target_index - 1, 0, -1)`), the rightward
// walk stops after len(fwd_primer_order) names for either strand, allowed_mismatches is 0, and an expected-name index of -1
// wraps to the last entry as a Python list index does.
#pragma once
#include "lsw_common.cuh"
#include "../../include/lsw.h"

namespace lsw {

// identity of a hit as the reference computes it: swalign path m / (m + x) (lamprey.py:566 via swalign), skbio path
// matches / len(primer) (lamprey.py:585); qlen = null selects the first.
LSW_HD double hit_identity(const lsw_hit& h, const int32_t* qlen) {
    return qlen ? (double)h.matches / (double)qlen[h.query] : (double)h.matches / (double)(h.matches + h.mismatches);
}

// removeOverlappingPrimerAlignments for the hits [lo, hi) of one read: a hit that starts more than `allowed` bases before the
// end of the current survivor competes with it and the strictly higher identity wins.  Survivors get reserved = 1.
LSW_HD void prune_overlaps_read(lsw_hit* hits, long long lo, long long hi, int allowed, const int32_t* qlen = nullptr) {
    if (lo >= hi) return;
    long long cur = lo;
    for (long long j = lo + 1; j < hi; j++) {
        if (hits[j].start < hits[cur].end - allowed) {
            if (hit_identity(hits[j], qlen) > hit_identity(hits[cur], qlen)) cur = j;
        } else {
            hits[cur].reserved = 1;
            cur = j;
        }
    }
    hits[cur].reserved = 1;
}

struct ChainOrders {
    int n_fwd, n_rev;
    int fwd[LSW_MAX_ORDER], rev[LSW_MAX_ORDER];   // query index of every name of fwd_primer_order / rev_primer_order (-1: not a schema sequence)
    int t_query, tc_query;                        // query indices of 'T' and 'Tc'
    int fwd_t, rev_tc;                            // fwd_primer_order.index('T'), rev_primer_order.index('Tc')
};

LSW_HD int order_at(const int* order, int n, int k) {      // Python list indexing: -1 is the last entry; out of range: no name
    if (k < 0) k += n;
    return (k >= 0 && k < n) ? order[k] : -2;
}

// The surviving hits (reserved == 1) of one read are `alignments` of lamprey.py:1536-1589.  For every target hit among them:
// extractAmpliconAroundTarget + the seq_end == -1 rule.  out may be null (count only).  Returns the number of targets.
LSW_HD int chain_subreads_read(const lsw_hit* hits, long long lo, long long hi, int read_len, const ChainOrders& C, lsw_subread* out) {
    long long first = -1;
    for (long long j = lo; j < hi; j++) if (hits[j].reserved == 1) { first = j; break; }
    if (first < 0) return 0;
    int n = 0;
    for (long long p = first; p < hi; p++) {
        if (hits[p].reserved != 1) continue;
        const int q = hits[p].query;
        if (q != C.t_query && q != C.tc_query) continue;
        if (out) {
            const bool fwd = q == C.t_query;
            const int* order = fwd ? C.fwd : C.rev;
            const int n_order = fwd ? C.n_fwd : C.n_rev;
            // to the right
            int amplicon_end = -1;
            int counter = (fwd ? C.fwd_t : C.rev_tc) + 1;
            long long prev = p;
            for (long long i = p + 1; i < hi; i++) {
                if (hits[i].reserved != 1) continue;
                if (hits[i].query != order_at(order, n_order, counter)) { amplicon_end = hits[prev].end; break; }
                counter++;
                if (counter == C.n_fwd) break;
                prev = i;
            }
            // to the left (never looks at the first survivor)
            int amplicon_start = 0;
            counter = (fwd ? C.fwd_t : C.rev_tc) - 1;
            long long next = p;
            for (long long i = p - 1; i > first; i--) {
                if (hits[i].reserved != 1) continue;
                if (hits[i].query != order_at(order, n_order, counter)) { amplicon_start = hits[next].start; break; }
                counter--;
                if (counter < 0) break;
                next = i;
            }
            lsw_subread s;
            s.read = hits[p].read; s.start = amplicon_start; s.end = amplicon_end == -1 ? read_len - 1 : amplicon_end;
            s.target_idx = n; s.direction = fwd ? 0 : 1; s.hit = (int32_t)(p - lo);
            out[n] = s;
        }
        n++;
    }
    return n;
}

}  // namespace lsw
