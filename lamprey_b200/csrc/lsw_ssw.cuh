// lsw_ssw.cuh -- the DEFAULT seeding path of LAMPrey (no --swalign): scikit-bio's StripedSmithWaterman.
//
//   /root/reference/lamprey.py:572-593  alignSequence_skbio(seq, primer, name):
//        StripedSmithWaterman(primer)(str(seq)) -> target_begin, target_end_optimal + 1, identity = matches / len(primer)
//   /root/reference/lamprey.py:609-610  findPrimerAlignments takes this branch unless args.swalign
//
// scikit-bio wraps Mengyao Zhao's SSW library (ssw.c); neither is in /root/reference.  What is computed here is the
// algorithm of ssw.c's ssw_align with skbio's default flag (begin positions and CIGAR), function by function:
//   1. forward   affine-gap local alignment of the primer (rows, in registers) against the read slice (columns, streamed):
//                H = max(0, D + w, E, F), E/F open with gapO and extend with gapE; the end is the FIRST column whose maximum
//                is strictly greater than everything before it and, in that column, the SMALLEST row holding the maximum
//                (the opposite of swalign's "last cell wins");
//   2. reverse   the same scan leftwards from that end over the reversed prefixes, stopped at the first column whose maximum
//                equals the score: target_begin and query_begin;
//   3. banded    banded_sw on the box between begin and end: direction codes with SSW's preferences (diagonal on ties, gap
//                open over gap extension on ties), band doubled until the score is reproduced, traceback -> CIGAR -> matches.
// One THREAD owns one task and runs all three steps, then the hit test and the left/right split of
// lamprey.py:612-640.  This is the first CUDA version of this row: scalar int32 cells (no s16x2 packing, no reuse of the
// round-0 matrix), all primer rows in registers; the recursion driver, task lists, hit records and sorting are shared with
// the swalign path (lsw.cu).
#pragma once
#include "lsw_common.cuh"
#include "lsw_trace.cuh"

namespace lsw {

struct SswParams {
    const uint8_t*  codes;        // read codes (0..3, CODE_OTHER = "N": scores 0 against everything, skbio _build_match_matrix)
    const int64_t*  roff;
    TaskArrays      t;
    const uint8_t*  qcodes;       // [n_seq][QSTRIDE] 0..3, padding 0xFF
    const int32_t*  qlen;
    const int64_t*  seg;          // [n_seq + 1] tasks of sequence s are [seg[s], seg[s+1])
    const int32_t*  class_seqs;   // sequences of this launch's register-row class
    int             n_class_seqs;
    int32_t match, mismatch, gap_open, gap_extend;   // match > 0, mismatch <= 0 (scores); gap penalties positive
    int32_t mode;                 // 0 = find_all (hit test, children), 1 = align_tasks (result records, CIGAR)
    double  thr;
    int32_t round;
    void*   hits; int64_t hits_cap; unsigned long long* nhits;
    uint8_t* child_flag; TaskArrays child;
    void*   results; uint32_t* cigar; int64_t cigar_cap; unsigned long long* ncigar;
    uint32_t* fallback;           // tasks whose band outgrew this launch's tier
    unsigned long long* nfallback;
    const uint32_t* fb_in;        // tier 2: the tasks to redo (null: every task of the class)
    const unsigned long long* fb_in_cnt;
    unsigned long long* counters; // [0] += tasks, [1] += cells
    int32_t* errflag;
    // with the packed K1 of lsw_affine.cuh in front: its per-task result and the candidates it queued (null: the scalar forward scan)
    const int2* res;
    const CandRec* cand_rec;
    const unsigned long long* ncand;
};

template <int PW>
struct PathBuf {                                             // 2 bits per op, in walk order (alignment end first)
    int nops;
    uint32_t ops[PW];
    LSW_HD void clear() { nops = 0; for (int i = 0; i < PW; i++) ops[i] = 0; }
    LSW_HD bool push(int op) {
        if (nops >= PW * 16) return false;
        ops[nops >> 4] |= (uint32_t)op << (2 * (nops & 15));
        nops++;
        return true;
    }
    LSW_HD int get(int i) const { return (int)((ops[i >> 4] >> (2 * (i & 15))) & 3u); }
};

template <int PW>
struct SswOutT {
    int score, ref_begin, ref_end, read_begin, read_end;     // 0-based, inclusive (-1: nothing aligned)
    int matches, others, nrun;
    PathBuf<PW> po;
};
using SswOut = SswOutT<MAX_PATH_WORDS>;                      // primer rows: a path has < 63 + 190 ops

template <int NR>
struct SswLane {
    static_assert(NR % 4 == 0 && NR <= 64, "rows in registers");

    // skbio's matrix: "N" (any non-ACGT base) scores 0 against everything, in the target and -- stage 2, where the query is a
    // sub-read -- in the query.  (A padding row of the register kernels carries 0xFF: never equal, plain mismatch.)
    LSW_HD static int sub(int qb, int bc, int match, int mismatch) {
        return (bc >= CODE_OTHER || qb == CODE_OTHER) ? 0 : (qb == bc ? match : mismatch);
    }

    // steps 1 and 2.  ref[c]: code of slice column c.  false: nothing aligned (score 0).
    LSW_HD static bool ends(const SswParams& P, const uint8_t* ref, int len, const uint32_t* qw, int Q, SswOut& o) {
        int H[NR], E[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) { H[r] = 0; E[r] = 0; }
        const int go = P.gap_open, ge = P.gap_extend, M = P.match, X = P.mismatch;
        int best = 0, end_ref = -1, end_read = 0;
        for (int c = 0; c < len; c++) {
            const int bc = ref[c];
            int diag = 0, F = 0, colmax = 0, colrow = 0;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int qb = (int)((qw[r >> 2] >> (8 * (r & 3))) & 0xFFu);
                int h = diag + sub(qb, bc, M, X);
                h = h > 0 ? h : 0;
                h = h > E[r] ? h : E[r];
                h = h > F ? h : F;
                diag = H[r]; H[r] = h;
                if (r < Q && h > colmax) { colmax = h; colrow = r; }        // first (smallest) row of the column's maximum
                int hg = h - go; hg = hg > 0 ? hg : 0;
                const int e = E[r] - ge; E[r] = e > hg ? e : hg;
                const int f = F - ge; F = f > hg ? f : hg;
            }
            if (colmax > best) { best = colmax; end_ref = c; end_read = colrow; }  // strictly greater: the first column wins
        }
        o.score = best; o.ref_end = end_ref; o.read_end = end_read;
        o.ref_begin = -1; o.read_begin = -1;
        if (best <= 0) return false;
        return begin_from_end(P, ref, qw, o);
    }

    // step 2 alone: o.score, o.ref_end, o.read_end are known (from the scan above, or from the packed K1 of lsw_affine.cuh)
    LSW_HD static bool begin_from_end(const SswParams& P, const uint8_t* ref, const uint32_t* qw, SswOut& o) {
        int H[NR], E[NR];
        const int go = P.gap_open, ge = P.gap_extend, M = P.match, X = P.mismatch;
        const int best = o.score, end_ref = o.ref_end, end_read = o.read_end;
        o.ref_begin = -1; o.read_begin = -1;
        // reverse scan from the end cell: rows end_read .. 0, columns end_ref .. 0, until a column's maximum equals the score
#pragma unroll
        for (int r = 0; r < NR; r++) { H[r] = 0; E[r] = 0; }
        int best2 = 0;
        for (int c = end_ref; c >= 0; c--) {
            const int bc = ref[c];
            int diag = 0, F = 0, colmax = 0, colrow = end_read;
#pragma unroll
            for (int r = NR - 1; r >= 0; r--) {
                if (r <= end_read) {
                    const int qb = (int)((qw[r >> 2] >> (8 * (r & 3))) & 0xFFu);
                    int h = diag + sub(qb, bc, M, X);
                    h = h > 0 ? h : 0;
                    h = h > E[r] ? h : E[r];
                    h = h > F ? h : F;
                    diag = H[r]; H[r] = h;
                    if (h > colmax) { colmax = h; colrow = r; }               // smallest index of the REVERSED read = largest row
                    int hg = h - go; hg = hg > 0 ? hg : 0;
                    const int e = E[r] - ge; E[r] = e > hg ? e : hg;
                    const int f = F - ge; F = f > hg ? f : hg;
                }
            }
            if (colmax > best2) { best2 = colmax; o.ref_begin = c; o.read_begin = colrow; }
            if (colmax == best) break;
        }
        return o.ref_begin >= 0;
    }
};

// step 3: ssw.c's banded_sw, statement by statement, with a band of at most BWMAX.
// dir: one byte per band cell: bit 0 = E came from H (code 3, else 2), bit 1 = F came from H (code 5, else 4), bits 2-4 = the
// H direction (1 diagonal, 2/3 E, 4/5 F).  Returns 1 ok, 0 the band outgrew BWMAX (next tier), -1 the box cannot reproduce the
// score / the traceback left the matrix (reported, never guessed).
template <int BWMAX>
struct SswBand {
    static constexpr int WMAX = 2 * BWMAX + 4;
    static constexpr int WD = 2 * BWMAX + 1;
    static constexpr int DIR_BYTES = WD * MAXQ;

    LSW_HD static int xoff(int i, int w) { const int x = i - w; return x > 0 ? x : 0; }

    template <class OUT>
    LSW_HD static int run(const SswParams& P, const uint8_t* ref, const uint8_t* qc, OUT& o, uint8_t* dir, long long dir_cap = DIR_BYTES) {
        const int refLen = o.ref_end - o.ref_begin + 1, readLen = o.read_end - o.read_begin + 1;
        const uint8_t* rf = ref + o.ref_begin;
        const uint8_t* rd = qc + o.read_begin;
        const int go = P.gap_open, ge = P.gap_extend;
        int band = (refLen > readLen ? refLen - readLen : readLen - refLen) + 1;
        int16_t h_b[WMAX], e_b[WMAX], h_c[WMAX];
        int width_d = 0, mx = 0;
        for (;;) {
            if (band > BWMAX || (long long)(band * 2 + 1) * readLen > dir_cap) return 0;
            const int width = band * 2 + 3;
            width_d = band * 2 + 1;
            for (int j = 0; j < width; j++) { h_b[j] = 0; e_b[j] = 0; h_c[j] = 0; }
            for (int i = 0; i < readLen; i++) {
                int beg = 0, end = refLen - 1, u = 0;
                int j = i - band; beg = beg > j ? beg : j;
                j = i + band; end = end < j ? end : j;
                const int edge = end + 1 < width - 1 ? end + 1 : width - 2;
                int f = 0;
                h_b[0] = 0; e_b[0] = 0; h_b[edge] = 0; e_b[edge] = 0; h_c[0] = 0;
                uint8_t* line = dir + (long long)width_d * i;
                const int x0 = xoff(i, band), x1 = xoff(i - 1, band);
                const int qb = rd[i];
                for (j = beg; j <= end; j++) {
                    u = j - x0 + 1;
                    const int e = j - x1 + 1, b = j - 1 - x0 + 1, d = j - 1 - x1 + 1;
                    int t1 = i == 0 ? -go : (int)h_b[e] - go;
                    int t2 = i == 0 ? -ge : (int)e_b[e] - ge;
                    const int ev = t1 > t2 ? t1 : t2;
                    e_b[u] = (int16_t)ev;
                    const int de = t1 > t2 ? 3 : 2;
                    t1 = (int)h_c[b] - go;
                    t2 = f - ge;
                    f = t1 > t2 ? t1 : t2;
                    const int df = t1 > t2 ? 5 : 4;
                    const int e1 = ev > 0 ? ev : 0, f1 = f > 0 ? f : 0;
                    t1 = e1 > f1 ? e1 : f1;
                    t2 = (int)h_b[d] + SswLane<16>::sub(qb, rf[j], P.match, P.mismatch);
                    const int hv = t1 > t2 ? t1 : t2;
                    h_c[u] = (int16_t)hv;
                    if (hv > mx) mx = hv;
                    const int dh = t1 <= t2 ? 1 : (e1 > f1 ? de : df);
                    line[j - x0] = (uint8_t)((de == 3 ? 1 : 0) | (df == 5 ? 2 : 0) | (dh << 2));
                }
                for (j = 1; j <= u; j++) h_b[j] = h_c[j];
            }
            if (mx >= o.score) break;
            if (band > refLen + readLen) return -1;
            band *= 2;
        }
        // trace back
        int i = readLen - 1, j = refLen - 1, state = 2, last = 0;
        o.po.clear();
        o.nrun = 0;
        bool ok = true;
        while (i > 0) {
            const int x = j - xoff(i, band);
            if (j < 0 || x < 0 || x >= width_d) return -1;
            const int cell = dir[(long long)width_d * i + x];
            const int code = state == 2 ? (cell >> 2) : state == 0 ? ((cell & 1) ? 3 : 2) : ((cell & 2) ? 5 : 4);
            int op;
            if (code == 1) { --i; --j; state = 2; op = OPC_M; }
            else if (code == 2) { --i; state = 0; op = OPC_I; }
            else if (code == 3) { --i; state = 2; op = OPC_I; }
            else if (code == 4) { --j; state = 1; op = OPC_D; }
            else if (code == 5) { --j; state = 2; op = OPC_D; }
            else return -1;
            if (op != last) { o.nrun++; last = op; }
            ok = ok && o.po.push(op);
        }
        if (last != OPC_M) o.nrun++;             // the cell of row 0 always closes the path with one 'M'
        ok = ok && o.po.push(OPC_M);
        if (!ok) return -1;
        // lamprey.py:579-583 on skbio's aligned sequences: walk the CIGAR forward from (target_begin, query_begin)
        int ti = 0, qi = 0, m = 0, others = 0;
        for (int k = o.po.nops - 1; k >= 0; k--) {
            const int op = o.po.get(k);
            if (op == OPC_M) {
                const int a = ti < refLen ? rf[ti] : CODE_OTHER, b2 = qi < readLen ? rd[qi] : 0xFF;
                if (a == b2 && a < CODE_OTHER) m++; else others++;
                ti++; qi++;
            } else if (op == OPC_I) { qi++; others++; }
            else { ti++; others++; }
        }
        o.matches = m; o.others = others;
        return 1;
    }
};

// One task, all steps.  Returns 1 done, 0 next tier (band), -1 internal inconsistency.
template <int NR, int BWMAX>
LSW_HD int ssw_task(const SswParams& P, int t, uint8_t* dir, bool have_end = false, int k_score = 0, int k_ref_end = 0,
                    int k_read_end = 0) {
    const int rd = P.t.read[t], a = P.t.a[t], b = P.t.b[t], q = P.t.seq[t];
    const int Q = P.qlen[q];
    const uint8_t* ref = P.codes + P.roff[rd] + a;
    const uint8_t* qc = P.qcodes + (size_t)q * QSTRIDE;
    uint32_t qw[NR / 4];
#pragma unroll
    for (int i = 0; i < NR / 4; i++) qw[i] = reinterpret_cast<const uint32_t*>(qc)[i];
    SswOut o;
    o.matches = 0; o.others = 0; o.nrun = 0; o.po.clear();
    bool aligned;
    if (have_end) {                      // K1 (lsw_affine.cuh) already found the score and the end cell
        o.score = k_score; o.ref_end = k_ref_end; o.read_end = k_read_end; o.ref_begin = -1; o.read_begin = -1;
        aligned = k_score > 0 && SswLane<NR>::begin_from_end(P, ref, qw, o);
        if (k_score > 0 && !aligned) return -1;
    } else {
        aligned = SswLane<NR>::ends(P, ref, b - a, qw, Q, o);
    }
    int rc = 1;
    if (aligned) {
        // find_all: a hit needs matches >= thr * Q, and matches <= query_end - query_begin + 1 <= span of the box; skip the
        // banded pass when even a perfect box cannot pass (exact: both are necessary conditions of lamprey.py:616-617)
        const int ql = o.read_end - o.read_begin + 1, rl = o.ref_end - o.ref_begin + 1;
        const bool hopeless = P.mode == 0 && (!((double)ql / (double)Q >= P.thr) || !((double)rl >= (double)Q * P.thr));
        if (!hopeless) {
            rc = SswBand<BWMAX>::run(P, ref, qc, o, dir);
            if (rc != 1) return rc;
        } else { o.matches = 0; }
    }
    if (P.mode == 1) {
        lsw_result* out = reinterpret_cast<lsw_result*>(P.results) + P.t.orig[t];
        out->score = o.score;
        out->q_pos = aligned ? o.read_begin : 0; out->q_end = aligned ? o.read_end + 1 : 0;
        out->r_pos = aligned ? o.ref_begin : -1; out->r_end = aligned ? o.ref_end + 1 : 0;
        out->matches = o.matches; out->mismatches = o.others;
        out->n_cigar = aligned ? o.nrun : 0;
        long long off = 0;
        if (P.cigar && aligned && o.nrun > 0) {
            off = (long long)atomic_add_ull(P.ncigar, (unsigned long long)o.nrun);
            if (off + o.nrun <= P.cigar_cap) {
                int k = 0, last = 0, cnt = 0;
                for (int i = o.po.nops - 1; i >= 0; i--) {
                    const int op = o.po.get(i);
                    if (op != last) { if (last) P.cigar[off + k++] = ((uint32_t)cnt << 2) | (uint32_t)last; last = op; cnt = 0; }
                    cnt++;
                }
                if (last) P.cigar[off + k++] = ((uint32_t)cnt << 2) | (uint32_t)last;
            }
        }
        out->cigar_off = off;
        return 1;
    }
    if (!aligned) return 1;
    const int start = o.ref_begin, end = o.ref_end + 1;
    const double identity = (double)o.matches / (double)Q;                      // float(matches)/float(len(primer))
    if (identity >= P.thr && (double)(end - start) >= (double)Q * P.thr) {
        const unsigned long long hi = append_slot(P.nhits);
        if ((long long)hi < P.hits_cap) {
            lsw_hit* h = reinterpret_cast<lsw_hit*>(P.hits) + hi;
            h->read = rd; h->start = a + start; h->end = a + end;
            h->query = (int16_t)q; h->depth = (int16_t)P.round;
            h->matches = (int16_t)o.matches; h->mismatches = (int16_t)o.others;
            h->score = (int16_t)o.score; h->q_pos = (int16_t)o.read_begin; h->q_end = (int16_t)(o.read_end + 1);
            h->reserved = 0; h->reserved2 = 0;
        }
        if (start >= Q) {
            const int64_t j = 2 * (int64_t)t;
            P.child.read[j] = rd; P.child.a[j] = a; P.child.b[j] = a + start; P.child.seq[j] = (int16_t)q;
            P.child_flag[j] = 1;
        }
        if (b - (a + end) >= Q) {
            const int64_t j = 2 * (int64_t)t + 1;
            P.child.read[j] = rd; P.child.a[j] = a + end; P.child.b[j] = b; P.child.seq[j] = (int16_t)q;
            P.child_flag[j] = 1;
        }
    }
    return 1;
}

// ---- stage 2 of processLamplicon: StripedSmithWaterman(sub-read)(target reference) ---------------------------------------
//   /root/reference/lamprey.py:1637-1658: every candidate sub-read (reverse-complemented for 'rev' targets) is the QUERY, the
//   target reference string (279 / 627 bp for the shipped targets) the target.  A query of any length does not fit the register
//   rows above: this path keeps (H, E) of every query row in a global scratch, laid out [row][thread] so that the threads of a
//   warp, which walk the rows in step, touch consecutive words; the columns are the reference.  Same three steps, same rules.
constexpr int SSWG_BW = 510;             // band limit (3 x 1024 int16 band rows in local memory)
constexpr int SSWG_PATH_WORDS = 512;     // up to 8192 CIGAR columns

struct SswGenParams {
    const uint8_t* codes; const int64_t* roff;
    const lsw_pair* pairs; int64_t n;
    const uint8_t* tcodes; int32_t tlen;
    int32_t match, mismatch, gap_open, gap_extend;
    uint32_t* he;            // [row][slot]: H | E << 16 of the current column
    uint8_t* qbuf;           // [slot][qcap]: the query's codes (reverse-complemented if asked)
    int32_t qcap;
    uint8_t* dir; long long dir_cap;     // [slot][dir_cap]
    lsw_result* results; uint32_t* cigar; int64_t cigar_cap; unsigned long long* ncigar;
    int32_t* errflag;
};

// One sub-read.  Returns 1 done, -1 band / path / scratch limits exceeded or an internal inconsistency (flagged, never guessed).
LSW_HD int ssw_generic_task(const SswGenParams& G, int64_t k, int slot, int nslots) {
    const lsw_pair pr = G.pairs[k];
    const int Lq = pr.end - pr.start;
    const uint8_t* src = G.codes + G.roff[pr.read] + pr.start;
    uint8_t* qb = G.qbuf + (size_t)slot * G.qcap;
    for (int q = 0; q < Lq; q++) {
        const int c = pr.revcomp ? src[Lq - 1 - q] : src[q];
        qb[q] = (uint8_t)((pr.revcomp && c < CODE_OTHER) ? 3 - c : c);      // A<->T, C<->G; anything else stays "N"
    }
    lsw_result* out = G.results + k;
    out->score = 0; out->q_pos = 0; out->q_end = 0; out->r_pos = -1; out->r_end = 0; out->matches = 0; out->mismatches = 0;
    out->n_cigar = 0; out->cigar_off = 0;
    const int go = G.gap_open, ge = G.gap_extend;
    uint32_t* he = G.he + slot;
    auto scan = [&](int c0, int c1, int step, int r0 /* first row */, int rstep, int nrows, int stop, int& best, int& bref, int& bread) {
        for (int q = 0; q < nrows; q++) he[(size_t)(r0 + q * rstep) * nslots] = 0u;
        best = 0; bref = -1; bread = r0;
        for (int c = c0; c != c1; c += step) {
            const int tc = G.tcodes[c];
            int diag = 0, F = 0, colmax = 0, colrow = r0;
            for (int q = 0; q < nrows; q++) {
                const int r = r0 + q * rstep;
                const uint32_t v = he[(size_t)r * nslots];
                const int hp = (int)(int16_t)(v & 0xFFFFu), E = (int)(int16_t)(v >> 16);
                int h = diag + SswLane<16>::sub(qb[r], tc, G.match, G.mismatch);
                h = h > 0 ? h : 0;
                h = h > E ? h : E;
                h = h > F ? h : F;
                diag = hp;
                if (h > colmax) { colmax = h; colrow = r; }
                int hg = h - go; hg = hg > 0 ? hg : 0;
                const int e = E - ge, f = F - ge;
                const int en = e > hg ? e : hg;
                F = f > hg ? f : hg;
                he[(size_t)r * nslots] = ((uint32_t)h & 0xFFFFu) | ((uint32_t)en << 16);
            }
            if (colmax > best) { best = colmax; bref = c; bread = colrow; }
            if (colmax == stop) break;
        }
    };
    if (Lq <= 0 || G.tlen <= 0) return 1;
    SswOutT<SSWG_PATH_WORDS> o;
    o.matches = 0; o.others = 0; o.nrun = 0;
    scan(0, G.tlen, 1, 0, 1, Lq, -1, o.score, o.ref_end, o.read_end);
    out->score = o.score;
    if (o.score <= 0 || o.score > 32000) return o.score > 32000 ? -1 : 1;
    int best2;
    scan(o.ref_end, -1, -1, o.read_end, -1, o.read_end + 1, o.score, best2, o.ref_begin, o.read_begin);
    if (o.ref_begin < 0) return -1;
    SswParams P;
    P.match = G.match; P.mismatch = G.mismatch; P.gap_open = go; P.gap_extend = ge;
    const int rc = SswBand<SSWG_BW>::run(P, G.tcodes, qb, o, G.dir + (size_t)slot * (size_t)G.dir_cap, G.dir_cap);
    if (rc != 1) return -1;
    out->q_pos = o.read_begin; out->q_end = o.read_end + 1;
    out->r_pos = o.ref_begin; out->r_end = o.ref_end + 1;
    out->matches = o.matches; out->mismatches = o.others; out->n_cigar = o.nrun;
    if (G.cigar && o.nrun > 0) {
        const long long off = (long long)atomic_add_ull(G.ncigar, (unsigned long long)o.nrun);
        out->cigar_off = off;
        if (off + o.nrun <= G.cigar_cap) {
            int kk = 0, last = 0, cnt = 0;
            for (int i = o.po.nops - 1; i >= 0; i--) {
                const int op = o.po.get(i);
                if (op != last) { if (last) G.cigar[off + kk++] = ((uint32_t)cnt << 2) | (uint32_t)last; last = op; cnt = 0; }
                cnt++;
            }
            if (last) G.cigar[off + kk++] = ((uint32_t)cnt << 2) | (uint32_t)last;
        }
    }
    return 1;
}

constexpr int SSW_THREADS = 128;
constexpr int SSW_BW1 = 8, SSW_BW2 = 128;     // band limits of the two tiers (banded_sw doubles its band until the score is reproduced)

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128)
k_ssw_generic(const SswGenParams G, int nslots) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= nslots) return;
    for (int64_t k = slot; k < G.n; k += nslots)
        if (ssw_generic_task(G, k, slot, nslots) != 1) *G.errflag = 1;
}


// Tier 1: one thread per task of the class, band up to SSW_BW1 (direction bytes in local memory).
// Tier 2: the few tasks whose band grew further, band up to SSW_BW2 (direction bytes in a global scratch).
template <int NR, int BWMAX>
__global__ void __launch_bounds__(SSW_THREADS)
k_ssw_round(const SswParams P, uint8_t* scratch) {
    __shared__ long long s_cum[260];
    __shared__ unsigned long long s_tasks, s_cells;
    if (threadIdx.x == 0) {
        long long c = 0;
        for (int i = 0; i < P.n_class_seqs; i++) { s_cum[i] = c; const int q = P.class_seqs[i]; c += P.seg[q + 1] - P.seg[q]; }
        s_cum[P.n_class_seqs] = c;
        s_tasks = 0; s_cells = 0;
    }
    __syncthreads();
    const long long total = P.fb_in ? (long long)*P.fb_in_cnt : P.cand_rec ? (long long)*P.ncand : s_cum[P.n_class_seqs];
    unsigned long long tasks = 0, cells = 0;
    uint8_t local_dir[BWMAX <= 16 ? SswBand<BWMAX>::DIR_BYTES : 1];
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        int t;
        bool have_end = false;
        int k_score = 0, k_ref_end = 0, k_read_end = 0;
        if (P.fb_in) t = (int)P.fb_in[k];
        else if (P.cand_rec) t = P.cand_rec[k].t;        // the candidates K1 queued (any order)
        else {
            int lo = 0, hi = P.n_class_seqs;                        // sequence of the class that owns the k-th task
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_cum[mid] <= k) lo = mid; else hi = mid; }
            t = (int)(P.seg[P.class_seqs[lo]] + (k - s_cum[lo]));
        }
        uint8_t* dir = BWMAX <= 16 ? local_dir
                                   : scratch + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (size_t)SswBand<BWMAX>::DIR_BYTES;
        if (P.res) {                                        // K1's result of the task: score | (end row + 1) << 16, end column + 1
            const int2 rs = P.res[t];
            have_end = true; k_score = rs.x & 0xFFFF; k_read_end = (rs.x >> 16) - 1; k_ref_end = rs.y - 1;
        }
        const int rc = ssw_task<NR, BWMAX>(P, t, dir, have_end, k_score, k_ref_end, k_read_end);
        if (rc == 0) { const unsigned long long j = atomicAdd(P.nfallback, 1ull); P.fallback[j] = (uint32_t)t; }
        else {
            if (rc < 0) *P.errflag = 1;
            if (!P.res) {                                   // (with K1 in front the tasks and cells are counted there)
                tasks += 1;
                cells += (unsigned long long)(P.t.b[t] - P.t.a[t]) * (unsigned long long)P.qlen[P.t.seq[t]];
            }
        }
    }
    atomicAdd(&s_tasks, tasks); atomicAdd(&s_cells, cells);
    __syncthreads();
    if (threadIdx.x == 0 && s_tasks) { atomicAdd(&P.counters[0], s_tasks); atomicAdd(&P.counters[1], s_cells); }
}
#endif

}  // namespace lsw
