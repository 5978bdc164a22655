// lsw_emu.cu -- TEST INFRASTRUCTURE: runs the per-lane logic of the CUDA kernels on the host.
//
// The kernels' lane code (ScoreLane<NR> in lsw_score.cuh, trace_task<NR> in lsw_trace.cuh) is
// __host__ __device__.  This file drives it one lane at a time with host loops in place of the
// kernel launches, the scan and the scatter, so tests/test_emu.py can check the exact
// arithmetic (packed DPX recurrence, arg-max fold, task switching, windowed traceback, hit test,
// recursion rounds) against the oracle in the build container, which has no GPU.
// It is built into its own library (liblsw_emu.so) and is never loaded by the product path:
// lamprey_b200 only opens liblsw.so and fails loudly without a CUDA device.
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <algorithm>

#include "../../include/lsw.h"
#include "lsw_common.cuh"
#include "lsw_score.cuh"
#include "lsw_trace.cuh"
#include "lsw_band.cuh"
#include "lsw_reuse.cuh"
#include "lsw_ssw.cuh"
#include "lsw_affine.cuh"
#include "lsw_host.hpp"

using namespace lsw;

namespace {

struct Emu {
    int match, mismatch, g;
    int nseq;
    std::vector<int32_t> qlen, class_of;
    std::vector<uint8_t> qcodes;
    std::vector<uint8_t> codes;
    std::vector<int64_t> roff;
    std::vector<int32_t> rlen;
    // task list
    std::vector<int32_t> t_read, t_a, t_b, t_orig;
    std::vector<int16_t> t_seq;
    std::vector<int64_t> seg;
    std::vector<int2> res;
    std::vector<uint8_t> cand_ok;
    unsigned long long counters[4];
    unsigned long long fallbacks = 0;
    bool force_full = false;
    // reuse of the round-0 DP (lsw_reuse.cuh)
    bool reuse = false;
    int round = 0;
    std::vector<uint32_t> blockbest;
    std::vector<int64_t> bb_off, bb1_off;
    int64_t bb_total = 0, bb1_total = 0;
    std::vector<uint32_t> bb1;       // second-level block bests (always used by the emulation when reuse is on)
    unsigned long long comp_cells = 0;
    int tiers = 1;                   // 1: band tiers of lsw_band.cuh (what the library runs); 0: the round-1 global-scratch kernel
    unsigned long long tier2 = 0;    // candidates the 8-row band passed on to the 16-row band
};

bool setup(Emu& E, int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int nseq, const uint8_t* qconcat,
           const int32_t* qoffs, int match, int mismatch, int gap) {
    if (match <= 0 || mismatch >= 0 || gap >= 0) return false;
    if (mismatch - gap < -128) return false;          // 256 * (mismatch + |gap|) must fit an s16 half (same check as lsw_set_scoring)
    E.match = match; E.mismatch = mismatch; E.g = -gap; E.nseq = nseq;
    E.qlen.resize(nseq); E.class_of.resize(nseq);
    E.qcodes.assign((size_t)nseq * QSTRIDE, 0xFF);
    for (int s = 0; s < nseq; s++) {
        const int len = qoffs[s + 1] - qoffs[s];
        if (len < 1 || len > LSW_MAX_QUERY_LEN || match * len + E.g > 127) return false;
        E.qlen[s] = len;
        for (int j = 0; j < len; j++) {
            const uint8_t c = encode_base(qconcat[qoffs[s] + j]);
            if (c >= CODE_OTHER) return false;
            E.qcodes[(size_t)s * QSTRIDE + j] = c;
        }
        int c = 0;
        while (CLASS_NR[c] < len) c++;
        E.class_of[s] = c;
    }
    E.roff.assign((size_t)n_reads + 1, 0);
    E.rlen.assign((size_t)n_reads, 0);
    int64_t d = 256;
    for (int64_t i = 0; i < n_reads; i++) {
        const int64_t len = offs[i + 1] - offs[i];
        E.roff[i] = d; E.rlen[i] = (int32_t)len;
        d += (len + 15) & ~15ll;
    }
    E.roff[n_reads] = d;
    E.codes.assign((size_t)d + 256, CODE_OTHER);
    for (int64_t i = 0; i < n_reads; i++)
        for (int64_t j = 0; j < E.rlen[i]; j++) E.codes[E.roff[i] + j] = encode_base(ascii[offs[i] + j]);
    memset(E.counters, 0, sizeof E.counters);
    return true;
}

template <int NR, class Lane = ScoreLane<NR>>
void emu_score_seq(Emu& E, const ScoreParams& P, int q, std::vector<uint32_t>& cand) {
    static U4 keep_ge[17], keep_lt[17];
    for (int n = 0; n <= 16; n++) make_keep_masks(keep_ge, keep_lt, n);
    std::vector<U2> prof((size_t)NIDX * Lane::PROF_STRIDE);
    const int Q = E.qlen[q];
    Lane::build_profile(P, q, Q, reinterpret_cast<uint32_t*>(prof.data()), 0, 1);
    Lane L;
    L.n_tasks = L.n_cells = L.n_steps = L.n_comp = 0;
    L.init();
    const long long cnt = P.seg[q + 1] - P.seg[q], base = P.seg[q];
    long long head = 0;
    bool exhausted = false;
    for (;;) {
        for (int h = 0; h < 2; h++) {
            if (L.tid[h] >= 0 && L.rem[h] == 0) L.finish(P, h, q, Q, [&](int t, int, int, int) { cand.push_back((uint32_t)t); });
            if (L.tid[h] < 0 && !exhausted) {
                if (head < cnt) { const int t = (int)(base + head++); L.start(P, h, t, P.kt[t], Q); }
                else exhausted = true;
            }
        }
        if (!(L.tid[0] >= 0 || L.tid[1] >= 0)) break;
        L.run_block(P, prof.data(), keep_ge, keep_lt, Q, q);
    }
    E.counters[0] += L.n_tasks; E.counters[1] += L.n_cells; E.counters[2] += L.n_steps; E.comp_cells += L.n_comp;
}

void emu_score_class(Emu& E, const ScoreParams& P, int c, int q, std::vector<uint32_t>& cand) {
    switch (CLASS_NR[c]) {
        case 16: emu_score_seq<16>(E, P, q, cand); break;
        case 20: emu_score_seq<20>(E, P, q, cand); break;
        case 24: emu_score_seq<24>(E, P, q, cand); break;
        case 28: emu_score_seq<28>(E, P, q, cand); break;
        case 32: emu_score_seq<32>(E, P, q, cand); break;
        case 40: emu_score_seq<40>(E, P, q, cand); break;
        case 48: emu_score_seq<48>(E, P, q, cand); break;
        case 56: emu_score_seq<56>(E, P, q, cand); break;
        default: emu_score_seq<64>(E, P, q, cand); break;
    }
}

template <int NR, int NR2>
void emu_trace_seq(Emu& E, const TraceParams& T, const ScoreParams& P, int q, const std::vector<uint32_t>& cand) {
    using Lane = TraceLane<NR>;
    std::vector<U2> prof((size_t)NIDX * Lane::PROF_STRIDE);
    ScoreLane<NR>::build_profile(P, q, E.qlen[q], reinterpret_cast<uint32_t*>(prof.data()), 0, 1);
    Lane L;
    // the records K1's finish / the merge kernel write on the device
    std::vector<CandRec> recs(cand.size());
    for (size_t i = 0; i < cand.size(); i++) {
        const int t = (int)cand[i];
        CandRec& c = recs[i];
        c.p0 = T.roff[T.t.read[t]] + T.t.a[t]; c.len = T.t.b[t] - T.t.a[t]; c.rd = T.t.read[t]; c.a = T.t.a[t];
        c.res_x = T.res[t].x; c.col = T.res[t].y; c.t = t;
    }
    TraceParams TR = T;
    static uint8_t dtab[128];
    for (int k = 0; k < 128; k++) dtab[k] = decision_entry(k);
    TR.cand_rec = recs.data();
    BandLane<NR, 2> B2;
    std::vector<U2> prof_lex((size_t)NIDX * Lane::PROF_STRIDE);      // the band tiers fill with the lexicographic scores
    {
        ScoreParams PL = P;
        PL.s_match = BandLane<NR, 2>::lex_match(T.match, T.g); PL.s_mis = BandLane<NR, 2>::lex_mis(T.mismatch, T.g);
        ScoreLane<NR>::build_profile(PL, q, E.qlen[q], reinterpret_cast<uint32_t*>(prof_lex.data()), 0, 1);
    }
    std::vector<uint32_t> band2(BandLane<NR, 2>::WORDS_PER_LANE);
    std::vector<int> tier2;
    for (size_t i = 0; i < cand.size(); i += 2) {
        const int t0 = (int)cand[i], t1 = i + 1 < cand.size() ? (int)cand[i + 1] : -1;
        E.counters[3] += t1 >= 0 ? 2 : 1;
        if (E.force_full || !nibble_walk_ok(T.match, T.mismatch, T.g)) {
            E.fallbacks += t1 >= 0 ? 2 : 1;
            trace_full_task<NR2>(T, t0, 0);
            if (t1 >= 0) trace_full_task<NR2>(T, t1, 0);
        } else if (E.tiers == 0) {     // the round-1 kernel: whole window in a global scratch
            if (T.mode == 1)
                L.template run_pair<true>(TR, prof.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, (int)i, t1 >= 0 ? (int)i + 1 : -1, 0, [&](int t) { E.fallbacks++; trace_full_task<NR2>(T, t, 0); });
            else
                L.template run_pair<false>(TR, prof.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, (int)i, t1 >= 0 ? (int)i + 1 : -1, 0, [&](int t) { E.fallbacks++; trace_full_task<NR2>(T, t, 0); });
        } else {                       // what the library runs: 16-row band in shared memory (lsw_band.cuh) ...
            auto fb1 = [&](int ci, int) { tier2.push_back(ci); };
            if (T.mode == 1) B2.template run_pair<true>(TR, prof_lex.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, (int)i, t1 >= 0 ? (int)i + 1 : -1, band2.data(), fb1);
            else B2.template run_pair<false>(TR, prof_lex.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, (int)i, t1 >= 0 ? (int)i + 1 : -1, band2.data(), fb1);
        }
    }
    E.tier2 += tier2.size();
    for (size_t i = 0; i < tier2.size(); i += 2) {           // ... then the whole window in a global scratch, then the full window
        const int c0 = tier2[i], c1 = i + 1 < tier2.size() ? tier2[i + 1] : -1;
        auto fb2 = [&](int t) { E.fallbacks++; trace_full_task<NR2>(T, t, 0); };
        if (T.mode == 1) L.template run_pair<true>(TR, prof.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, c0, c1, 0, fb2);
        else L.template run_pair<false>(TR, prof.data(), T.qcodes + (size_t)q * QSTRIDE, dtab, q, c0, c1, 0, fb2);
    }
}

void emu_trace_class(Emu& E, const TraceParams& T, const ScoreParams& P, int c, int q, const std::vector<uint32_t>& cand) {
    switch (CLASS_NR[c]) {
        case 16: emu_trace_seq<16, 16>(E, T, P, q, cand); break;
        case 20: emu_trace_seq<20, 32>(E, T, P, q, cand); break;
        case 24: emu_trace_seq<24, 32>(E, T, P, q, cand); break;
        case 28: emu_trace_seq<28, 32>(E, T, P, q, cand); break;
        case 32: emu_trace_seq<32, 32>(E, T, P, q, cand); break;
        case 40: emu_trace_seq<40, 48>(E, T, P, q, cand); break;
        case 48: emu_trace_seq<48, 48>(E, T, P, q, cand); break;
        case 56: emu_trace_seq<56, 64>(E, T, P, q, cand); break;
        default: emu_trace_seq<64, 64>(E, T, P, q, cand); break;
    }
}

TaskArrays view(Emu& E) {
    TaskArrays t;
    t.read = E.t_read.data(); t.a = E.t_a.data(); t.b = E.t_b.data(); t.seq = E.t_seq.data(); t.orig = E.t_orig.data();
    t.skip = nullptr;
    return t;
}

// K1 + K2 over the current task list.  Tbase carries the mode-specific outputs.
// Round 0 (and every round when reuse is off, and align_tasks) runs K1 on the tasks themselves; later rounds
// run K1 on the segment list of lsw_reuse.cuh and merge per task.
void emu_round(Emu& E, TraceParams T) {
    const int64_t n = (int64_t)E.t_read.size();
    E.res.assign((size_t)n, make_int2(0, 0));
    ScoreParams P;
    memset(&P, 0, sizeof P);
    P.codes = E.codes.data(); P.roff = E.roff.data(); P.t = view(E); P.seg = E.seg.data();
    P.res = E.res.data(); P.qcodes = E.qcodes.data(); P.qlen = E.qlen.data();
    P.cand_ok = E.cand_ok.data();
    P.s_match = SC * (E.match + E.g); P.s_mis = SC * (E.mismatch + E.g); P.g = E.g; P.match = E.match;
    P.push_cands = 1; P.count_ref = 1; P.block_steps = BLOCK_STEPS;
    int qmax = 1;
    for (int s = 0; s < E.nseq; s++) qmax = std::max(qmax, (int)E.qlen[s]);
    // one lane (slot 0): fast layout = rows of TRACE_THREADS x 16 B per (column, unit pair), at most 5 pairs; full layout = 4 words per column
    std::vector<uint4> scratch((size_t)std::max(trace_window_full(qmax, E.match, E.g), (trace_window_fast_max(qmax, E.match, E.g) + 4) * 5 * TRACE_THREADS) + 4);
    T.codes = E.codes.data(); T.roff = E.roff.data(); T.t = view(E); T.res = E.res.data();
    T.qcodes = E.qcodes.data(); T.qlen = E.qlen.data(); T.cand_ok = E.cand_ok.data();
    T.match = E.match; T.mismatch = E.mismatch; T.g = E.g;
    T.scratch = reinterpret_cast<uint32_t*>(scratch.data()); T.scratch_stride = 1;

    std::vector<KTask> kt((size_t)n);
    for (int64_t t = 0; t < n; t++) { kt[t].p0 = E.roff[E.t_read[t]] + E.t_a[t]; kt[t].len = E.t_b[t] - E.t_a[t]; kt[t].skip = 0; }
    P.kt = kt.data();
    const bool segmented = E.reuse && E.round > 0;
    if (E.reuse && E.round == 0) { P.blockbest = E.blockbest.data(); P.bb_off = E.bb_off.data(); P.bb_total = E.bb_total; }
    if (!segmented) {
        for (int s = 0; s < E.nseq; s++) {
            if (E.seg[s + 1] == E.seg[s]) continue;
            std::vector<uint32_t> cand;
            emu_score_class(E, P, E.class_of[s], s, cand);
            emu_trace_class(E, T, P, E.class_of[s], s, cand);
        }
        return;
    }
    // plan -> scan -> fill -> K1 on segments -> merge -> K2
    ReuseParams R;
    memset(&R, 0, sizeof R);
    R.t = view(E); R.n = n; R.qlen = E.qlen.data(); R.match = E.match; R.g = E.g; R.enable = 1;
    R.blockbest = E.blockbest.data(); R.bb_off = E.bb_off.data(); R.bb_total = E.bb_total;
    std::vector<int32_t> wx;
    build_reuse_wx(E.nseq, E.qlen.data(), E.match, E.g, E.cand_ok, getenv("LSW_REUSE_FULL_HEAD") != nullptr, wx);
    R.wx = wx.data();
    if (E.round == 1 && E.bb1_total > 0) {           // built once, after round 0
        E.bb1.assign((size_t)E.bb1_total * E.nseq, 0u);
        const int64_t nr = (int64_t)E.bb_off.size() - 1;
        for (int q = 0; q < E.nseq; q++)
            for (int64_t rd = 0; rd < nr; rd++)
                for (int64_t G = 0; G < E.bb1_off[rd + 1] - E.bb1_off[rd]; G++)
                    E.bb1[(size_t)q * E.bb1_total + E.bb1_off[rd] + G] =
                        reduce_block_group(E.blockbest.data() + (size_t)q * E.bb_total + E.bb_off[rd],
                                           (int)(E.bb_off[rd + 1] - E.bb_off[rd]), (int)G, E.g);
    }
    if (E.bb1_total > 0) { R.bb1 = E.bb1.data(); R.bb1_off = E.bb1_off.data(); R.bb1_total = E.bb1_total; }
    std::vector<int32_t> segpos((size_t)n + 1, 0);
    for (int64_t t = 0; t < n; t++)
        segpos[t + 1] = segpos[t] + plan_of(R, t).nseg;
    const int64_t ns = segpos[n];
    std::vector<int2> res_s((size_t)ns, make_int2(0, 0));
    std::vector<int64_t> seg_s(E.nseq + 1);
    for (int s = 0; s <= E.nseq; s++) seg_s[s] = segpos[E.seg[s]];
    R.segpos = segpos.data();
    std::vector<KTask> s_kt((size_t)ns);
    R.s_kt = s_kt.data(); R.roff = E.roff.data();
    for (int64_t t = 0; t < n; t++) fill_segments(R, t);
    ScoreParams PS = P;
    PS.t = TaskArrays(); PS.kt = s_kt.data(); PS.seg = seg_s.data(); PS.res = res_s.data(); PS.push_cands = 0; PS.count_ref = 0; PS.block_steps = SEG_BLOCK;
    for (int s = 0; s < E.nseq; s++) {
        if (seg_s[s + 1] == seg_s[s]) continue;
        std::vector<uint32_t> none;
        emu_score_class(E, PS, E.class_of[s], s, none);
    }
    R.res_s = res_s.data(); R.res_t = E.res.data();
    R.blockbest = E.blockbest.data(); R.bb_off = E.bb_off.data(); R.bb_total = E.bb_total;
    R.cand_ok = E.cand_ok.data();
    for (int s = 0; s < E.nseq; s++) {
        std::vector<uint32_t> cand;
        for (int64_t t = E.seg[s]; t < E.seg[s + 1]; t++) {
            int sc, rw, cl;
            E.counters[0]++; E.counters[1] += (unsigned long long)(E.t_b[t] - E.t_a[t]) * E.qlen[s];
            if (combine_task(R, t, sc, rw, cl)) cand.push_back((uint32_t)t);
        }
        if (!cand.empty()) emu_trace_class(E, T, P, E.class_of[s], s, cand);
    }
}

}  // namespace

extern "C" {

int lsw_emu_find_all(int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int nseq, const uint8_t* qconcat,
                     const int32_t* qoffs, int match, int mismatch, int gap, double thr, lsw_hit* out, int64_t cap,
                     int64_t* n_out, int64_t* counters /* tasks, cells, steps, rounds, traced, fallbacks, computed cells, tier-2 candidates */, int force_full,
                     int allowed_overlap, int reuse) {
    Emu E;
    E.force_full = force_full == 1;          // force_full: 1 = every candidate through k_trace_full's logic, 2 = the round-1 global-scratch walker
    E.tiers = force_full == 2 ? 0 : 1;
    E.reuse = reuse != 0;
    if (!setup(E, n_reads, ascii, offs, nseq, qconcat, qoffs, match, mismatch, gap)) return LSW_E_UNSUPPORTED;
    build_cand_ok(nseq, E.qlen.data(), match, mismatch, E.g, thr, false, E.cand_ok);
    const int64_t n0 = n_reads * nseq;
    E.t_read.resize(n0); E.t_a.resize(n0); E.t_b.resize(n0); E.t_seq.resize(n0);
    E.seg.resize(nseq + 1);
    for (int s = 0; s <= nseq; s++) E.seg[s] = (int64_t)s * n_reads;
    for (int64_t i = 0; i < n0; i++) {
        const int s = (int)(i / n_reads); const int64_t r = i - (int64_t)s * n_reads;
        E.t_read[i] = (int32_t)r; E.t_a[i] = 0; E.t_b[i] = E.rlen[r]; E.t_seq[i] = (int16_t)s;
    }
    if (E.reuse) {
        E.bb_off.assign((size_t)n_reads + 1, 0);
        for (int64_t i = 0; i < n_reads; i++) E.bb_off[i + 1] = E.bb_off[i] + (E.rlen[i] + BB_BLOCK - 1) / BB_BLOCK;
        E.bb_total = E.bb_off[n_reads];
        E.bb1_off.assign((size_t)n_reads + 1, 0);
        for (int64_t i = 0; i < n_reads; i++) E.bb1_off[i + 1] = E.bb1_off[i] + (E.bb_off[i + 1] - E.bb_off[i] + BB_GROUP - 1) / BB_GROUP;
        E.bb1_total = E.bb1_off[n_reads];
        E.blockbest.assign((size_t)std::max<int64_t>(1, E.bb_total * nseq), 0u);
    }
    unsigned long long nhits = 0;
    int32_t errflag = 0;
    int round = 0;
    while (!E.t_read.empty()) {
        E.round = round;
        const int64_t n = (int64_t)E.t_read.size(), n2 = 2 * n;
        std::vector<int32_t> c_read(n2), c_a(n2), c_b(n2);
        std::vector<int16_t> c_seq(n2);
        std::vector<uint8_t> flag(n2 + 1, 0);
        TraceParams T;
        memset(&T, 0, sizeof T);
        T.mode = 0; T.thr = thr; T.round = round;
        T.hits = out; T.hits_cap = cap; T.nhits = &nhits;
        T.child_flag = flag.data();
        T.child.read = c_read.data(); T.child.a = c_a.data(); T.child.b = c_b.data(); T.child.seq = c_seq.data();
        T.errflag = &errflag;
        emu_round(E, T);
        // scan + scatter
        std::vector<int32_t> nr, na, nb; std::vector<int16_t> ns;
        std::vector<int64_t> nseg(nseq + 1, 0);
        std::vector<int32_t> pos(n2 + 1, 0);
        for (int64_t j = 0; j < n2; j++) pos[j + 1] = pos[j] + flag[j];
        for (int s = 0; s <= nseq; s++) nseg[s] = pos[std::min<int64_t>(2 * E.seg[s], n2)];
        for (int64_t j = 0; j < n2; j++)
            if (flag[j]) { nr.push_back(c_read[j]); na.push_back(c_a[j]); nb.push_back(c_b[j]); ns.push_back(c_seq[j]); }
        E.t_read.swap(nr); E.t_a.swap(na); E.t_b.swap(nb); E.t_seq.swap(ns); E.seg.swap(nseg);
        round++;
        if (round > 30000) return LSW_E_INTERNAL;      // as lsw_find_all: "recursion did not terminate" (lsw_hit.depth is 16 bits)
    }
#if defined(LSW_EMU_STATS)
    fprintf(stderr, "[emu] mean window columns %.1f; columns touched beyond max_row (histogram -20..40):", (double)g_ncols_sum / (double)(g_walks ? g_walks : 1));
    for (int i = 0; i < 61; i++) if (g_extra_hist[i]) fprintf(stderr, " %d:%llu", i - 20, g_extra_hist[i]);
    fprintf(stderr, "\n");
    for (int b = 1; b <= 2; b++)
        fprintf(stderr, "[emu] band BW=%d: pruned %llu walks %llu failed %llu (%.3f %%) iterations %.1f per walk\n", b, g_band_pruned[b], g_band_walks[b], g_band_fail[b],
                100.0 * (double)g_band_fail[b] / (double)(g_band_walks[b] ? g_band_walks[b] : 1), (double)g_band_iters[b] / (double)(g_band_walks[b] ? g_band_walks[b] : 1));
    for (int b = 1; b <= 2; b++)
        fprintf(stderr, "[emu] band BW=%d fail reasons: out-of-band %llu (low/path %llu low/nested %llu high/path %llu high/nested %llu) nibble %llu edge %llu budget %llu notie %llu depth %llu cert %llu\n", b,
                g_band_why[b][0], g_band_out[b][0][0], g_band_out[b][0][1], g_band_out[b][1][0], g_band_out[b][1][1], g_band_why[b][1], g_band_why[b][2], g_band_why[b][3], g_band_why[b][4], g_band_why[b][5], g_band_cert[b]);
    for (int b = 1; b <= 2; b++) {
        fprintf(stderr, "[emu] band BW=%d completed walks: -dmin hist:", b); for (int i = 0; i < 12; i++) fprintf(stderr, " %llu", g_dmin[b][i]);
        fprintf(stderr, "  dmax hist:"); for (int i = 0; i < 12; i++) fprintf(stderr, " %llu", g_dmax[b][i]);
        fprintf(stderr, "  exit row hist:"); for (int i = 0; i < 16; i++) fprintf(stderr, " %llu", g_drow[b][i]);
        fprintf(stderr, "\n");
    }
    fprintf(stderr, "[emu] walks %llu iters %llu (%.1f per walk) fill steps %llu (%.1f per pair)\n", g_walks, g_walk_iters, (double)g_walk_iters / (double)(g_walks ? g_walks : 1), g_fill_steps, 2.0 * g_fill_steps / (double)(g_walks ? g_walks : 1));
#endif
    if (errflag) return LSW_E_INTERNAL;
    if (n_out) *n_out = (int64_t)nhits;
    if (counters) { counters[0] = E.counters[0]; counters[1] = E.counters[1]; counters[2] = E.counters[2]; counters[3] = round; counters[4] = E.counters[3]; counters[5] = E.fallbacks; counters[6] = (int64_t)E.comp_cells; counters[7] = (int64_t)E.tier2; }
    if ((int64_t)nhits > cap) return LSW_E_OVERFLOW;
    // (read, start, query) order
    std::sort(out, out + nhits, [](const lsw_hit& x, const lsw_hit& y) {
        if (x.read != y.read) return x.read < y.read;
        if (x.start != y.start) return x.start < y.start;
        return x.query < y.query;
    });
    if (allowed_overlap >= 0) {
        long long lo = 0;
        while (lo < (long long)nhits) {
            long long hi = lo;
            while (hi < (long long)nhits && out[hi].read == out[lo].read) hi++;
            prune_overlaps_read(out, lo, hi, allowed_overlap);
            lo = hi;
        }
    }
    return LSW_OK;
}

int lsw_emu_align_tasks(int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int nseq, const uint8_t* qconcat,
                        const int32_t* qoffs, int match, int mismatch, int gap, int64_t n, const lsw_task* tasks,
                        lsw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used, int force_full,
                        int64_t* fallbacks) {
    Emu E;
    E.force_full = force_full == 1;          // force_full: 1 = every candidate through k_trace_full's logic, 2 = the round-1 global-scratch walker
    E.tiers = force_full == 2 ? 0 : 1;
    if (!setup(E, n_reads, ascii, offs, nseq, qconcat, qoffs, match, mismatch, gap)) return LSW_E_UNSUPPORTED;
    build_cand_ok(nseq, E.qlen.data(), match, mismatch, E.g, 0.0, true, E.cand_ok);
    E.seg.assign(nseq + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        const lsw_task& t = tasks[i];
        if (t.query < 0 || t.query >= nseq || t.read < 0 || t.read >= n_reads || t.start < 0 || t.end < t.start ||
            t.end > E.rlen[t.read]) return LSW_E_ARG;
        E.seg[t.query + 1]++;
    }
    for (int s = 0; s < nseq; s++) E.seg[s + 1] += E.seg[s];
    E.t_read.resize(n); E.t_a.resize(n); E.t_b.resize(n); E.t_seq.resize(n); E.t_orig.resize(n);
    std::vector<int64_t> fill(E.seg.begin(), E.seg.end() - 1);
    for (int64_t i = 0; i < n; i++) {
        const lsw_task& t = tasks[i];
        const int64_t d = fill[t.query]++;
        E.t_read[d] = t.read; E.t_a[d] = t.start; E.t_b[d] = t.end; E.t_seq[d] = (int16_t)t.query; E.t_orig[d] = (int32_t)i;
    }
    unsigned long long ncig = 0;
    int32_t errflag = 0;
    TraceParams T;
    memset(&T, 0, sizeof T);
    T.mode = 1; T.results = out; T.cigar = cigar_buf; T.cigar_cap = cigar_buf ? cigar_cap : 0; T.ncigar = &ncig;
    T.errflag = &errflag;
    if (n) emu_round(E, T);
    if (fallbacks) *fallbacks = (int64_t)E.fallbacks;
    if (errflag) return LSW_E_INTERNAL;
    if (cigar_used) *cigar_used = (int64_t)ncig;
    if (cigar_buf && (int64_t)ncig > cigar_cap) return LSW_E_OVERFLOW;
    return LSW_OK;
}


}  // extern "C"

// ---- the default (skbio / SSW) seeding path: lsw_ssw.cuh driven one task at a time --------------------------------------
namespace {
template <int NR>
int emu_ssw_one(const SswParams& P, int t, std::vector<uint8_t>& dir, unsigned long long* tier2) {
    dir.assign(SswBand<SSW_BW2>::DIR_BYTES, 0);
    bool have = false;
    int sc = 0, re = 0, rr = 0;
    if (P.res) { const int2 rs = P.res[t]; have = true; sc = rs.x & 0xFFFF; rr = (rs.x >> 16) - 1; re = rs.y - 1; }
    int rc = ssw_task<NR, SSW_BW1>(P, t, dir.data(), have, sc, re, rr);
    if (rc == 0) { if (tier2) (*tier2)++; rc = ssw_task<NR, SSW_BW2>(P, t, dir.data(), have, sc, re, rr); }
    return rc;
}
// the packed K1 of lsw_affine.cuh for every task of sequence q (tasks [seg[q], seg[q+1]) of the emulation's task list)
void emu_affine_seq(Emu& E, const ScoreParams& P, int q) {
    std::vector<uint32_t> cand;
    int c = 0;
    while (CLASS_NR[c] < E.qlen[q]) c++;
    switch (CLASS_NR[c]) {
        case 16: emu_score_seq<16, AffineLane<16>>(E, P, q, cand); break;
        case 20: emu_score_seq<20, AffineLane<20>>(E, P, q, cand); break;
        case 24: emu_score_seq<24, AffineLane<24>>(E, P, q, cand); break;
        case 28: emu_score_seq<28, AffineLane<28>>(E, P, q, cand); break;
        case 32: emu_score_seq<32, AffineLane<32>>(E, P, q, cand); break;
        case 40: emu_score_seq<40, AffineLane<40>>(E, P, q, cand); break;
        case 48: emu_score_seq<48, AffineLane<48>>(E, P, q, cand); break;
        case 56: emu_score_seq<56, AffineLane<56>>(E, P, q, cand); break;
        default: emu_score_seq<64, AffineLane<64>>(E, P, q, cand); break;
    }
}

// K1 (packed) over a task list grouped by sequence (E.seg): fills E.res
void emu_affine_round(Emu& E, int match, int mismatch, int gap_open, int gap_extend, std::vector<KTask>& kt, std::vector<uint8_t>& ok) {
    const int64_t n = (int64_t)E.t_read.size();
    E.res.assign((size_t)n, make_int2(0, 0));
    kt.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) { kt[i].p0 = E.roff[E.t_read[i]] + E.t_a[i]; kt[i].len = E.t_b[i] - E.t_a[i]; kt[i].skip = 0; }
    ok.assign((size_t)E.nseq * 128, 0);
    ScoreParams P;
    memset(&P, 0, sizeof P);
    P.codes = E.codes.data(); P.roff = E.roff.data();
    P.t.read = E.t_read.data(); P.t.a = E.t_a.data(); P.t.b = E.t_b.data(); P.t.seq = E.t_seq.data();
    P.kt = kt.data(); P.seg = E.seg.data(); P.res = E.res.data();
    P.qcodes = E.qcodes.data(); P.qlen = E.qlen.data(); P.cand_ok = ok.data();
    P.s_match = SC * match; P.s_mis = SC * mismatch; P.aff_open = gap_open; P.aff_extend = gap_extend;
    P.push_cands = 0; P.count_ref = 1; P.block_steps = BLOCK_STEPS;
    P.counters = E.counters;
    for (int s = 0; s < E.nseq; s++)
        if (E.seg[s + 1] > E.seg[s]) emu_affine_seq(E, P, s);
}

int emu_ssw_task(const SswParams& P, int t, int Q, std::vector<uint8_t>& dir, unsigned long long* tier2) {
    int c = 0;
    while (CLASS_NR[c] < Q) c++;
    switch (CLASS_NR[c]) {
        case 16: return emu_ssw_one<16>(P, t, dir, tier2);
        case 20: return emu_ssw_one<20>(P, t, dir, tier2);
        case 24: return emu_ssw_one<24>(P, t, dir, tier2);
        case 28: return emu_ssw_one<28>(P, t, dir, tier2);
        case 32: return emu_ssw_one<32>(P, t, dir, tier2);
        case 40: return emu_ssw_one<40>(P, t, dir, tier2);
        case 48: return emu_ssw_one<48>(P, t, dir, tier2);
        case 56: return emu_ssw_one<56>(P, t, dir, tier2);
        default: return emu_ssw_one<64>(P, t, dir, tier2);
    }
}
void ssw_base_params(Emu& E, SswParams& P, int match, int mismatch, int gap_open, int gap_extend) {
    memset(&P, 0, sizeof P);
    P.codes = E.codes.data(); P.roff = E.roff.data();
    P.qcodes = E.qcodes.data(); P.qlen = E.qlen.data();
    P.match = match; P.mismatch = mismatch; P.gap_open = gap_open; P.gap_extend = gap_extend;
}
}  // namespace

extern "C" {

int lsw_emu_ssw_find_all(int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int nseq, const uint8_t* qconcat,
                         const int32_t* qoffs, int match, int mismatch, int gap_open, int gap_extend, double thr, lsw_hit* out,
                         int64_t cap, int64_t* n_out, int64_t* counters /* tasks, cells, rounds, tier-2 tasks */,
                         int packed /* 1: forward scan by the packed K1 of lsw_affine.cuh (what the library runs); 0: scalar scan */) {
    Emu E;
    if (!setup(E, n_reads, ascii, offs, nseq, qconcat, qoffs, 2, -1, -1)) return LSW_E_UNSUPPORTED;
    const int64_t n0 = n_reads * nseq;
    E.t_read.resize(n0); E.t_a.resize(n0); E.t_b.resize(n0); E.t_seq.resize(n0);
    E.seg.resize(nseq + 1);
    for (int s = 0; s <= nseq; s++) E.seg[s] = (int64_t)s * n_reads;
    std::vector<KTask> kt;
    std::vector<uint8_t> okt;
    for (int s = 0; s < nseq; s++) if (match * E.qlen[s] > 127) packed = 0;      // outside the 8.8 cells: the scalar scan (as the library does)
    for (int64_t i = 0; i < n0; i++) {
        const int s = (int)(i / n_reads); const int64_t r = i - (int64_t)s * n_reads;
        E.t_read[i] = (int32_t)r; E.t_a[i] = 0; E.t_b[i] = E.rlen[r]; E.t_seq[i] = (int16_t)s;
    }
    unsigned long long nhits = 0, tasks = 0, cells = 0, tier2 = 0;
    int32_t errflag = 0;
    int round = 0;
    std::vector<uint8_t> dir;
    while (!E.t_read.empty()) {
        const int64_t n = (int64_t)E.t_read.size(), n2 = 2 * n;
        std::vector<int32_t> c_read(n2), c_a(n2), c_b(n2);
        std::vector<int16_t> c_seq(n2);
        std::vector<uint8_t> flag(n2 + 1, 0);
        SswParams P;
        ssw_base_params(E, P, match, mismatch, gap_open, gap_extend);
        P.t.read = E.t_read.data(); P.t.a = E.t_a.data(); P.t.b = E.t_b.data(); P.t.seq = E.t_seq.data();
        P.mode = 0; P.thr = thr; P.round = round;
        P.hits = out; P.hits_cap = cap; P.nhits = &nhits;
        P.child_flag = flag.data();
        P.child.read = c_read.data(); P.child.a = c_a.data(); P.child.b = c_b.data(); P.child.seq = c_seq.data();
        if (packed) { emu_affine_round(E, match, mismatch, gap_open, gap_extend, kt, okt); P.res = E.res.data(); }
        for (int64_t t = 0; t < n; t++) {
            tasks++; cells += (unsigned long long)(E.t_b[t] - E.t_a[t]) * E.qlen[E.t_seq[t]];
            if (emu_ssw_task(P, (int)t, E.qlen[E.t_seq[t]], dir, &tier2) != 1) errflag = 1;
        }
        std::vector<int32_t> nr, na, nb; std::vector<int16_t> ns;
        std::vector<int64_t> nseg(nseq + 1, 0);
        std::vector<int64_t> pos(n2 + 1, 0);
        for (int64_t j = 0; j < n2; j++) pos[j + 1] = pos[j] + flag[j];
        for (int s = 0; s <= nseq; s++) nseg[s] = pos[std::min<int64_t>(2 * E.seg[s], n2)];
        for (int64_t j = 0; j < n2; j++)
            if (flag[j]) { nr.push_back(c_read[j]); na.push_back(c_a[j]); nb.push_back(c_b[j]); ns.push_back(c_seq[j]); }
        E.t_read.swap(nr); E.t_a.swap(na); E.t_b.swap(nb); E.t_seq.swap(ns); E.seg.swap(nseg);
        round++;
        if (round > 30000) return LSW_E_INTERNAL;      // as lsw_find_all: "recursion did not terminate" (lsw_hit.depth is 16 bits)
    }
    if (errflag) return LSW_E_INTERNAL;
    if (n_out) *n_out = (int64_t)nhits;
    if (counters) { counters[0] = (int64_t)tasks; counters[1] = (int64_t)cells; counters[2] = round; counters[3] = (int64_t)tier2; }
    if ((int64_t)nhits > cap) return LSW_E_OVERFLOW;
    std::sort(out, out + nhits, [](const lsw_hit& x, const lsw_hit& y) {
        if (x.read != y.read) return x.read < y.read;
        if (x.start != y.start) return x.start < y.start;
        return x.query < y.query;
    });
    return LSW_OK;
}

int lsw_emu_ssw_align_tasks(int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int nseq, const uint8_t* qconcat,
                            const int32_t* qoffs, int match, int mismatch, int gap_open, int gap_extend, int64_t n,
                            const lsw_task* tasks, lsw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used,
                            int packed) {
    Emu E;
    if (!setup(E, n_reads, ascii, offs, nseq, qconcat, qoffs, 2, -1, -1)) return LSW_E_UNSUPPORTED;
    E.t_read.resize(n); E.t_a.resize(n); E.t_b.resize(n); E.t_seq.resize(n); E.t_orig.resize(n);
    E.seg.assign(nseq + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        const lsw_task& t = tasks[i];
        if (t.query < 0 || t.query >= nseq || t.read < 0 || t.read >= n_reads || t.start < 0 || t.end < t.start ||
            t.end > E.rlen[t.read]) return LSW_E_ARG;
        E.seg[t.query + 1]++;
    }
    for (int s = 0; s < nseq; s++) E.seg[s + 1] += E.seg[s];
    {
        std::vector<int64_t> fill(E.seg.begin(), E.seg.end() - 1);
        for (int64_t i = 0; i < n; i++) {
            const lsw_task& t = tasks[i];
            const int64_t d = fill[t.query]++;
            E.t_read[d] = t.read; E.t_a[d] = t.start; E.t_b[d] = t.end; E.t_seq[d] = (int16_t)t.query; E.t_orig[d] = (int32_t)i;
        }
    }
    std::vector<KTask> kt;
    std::vector<uint8_t> okt;
    for (int s = 0; s < nseq; s++) if (match * E.qlen[s] > 127) packed = 0;
    unsigned long long ncig = 0;
    SswParams P;
    ssw_base_params(E, P, match, mismatch, gap_open, gap_extend);
    P.t.read = E.t_read.data(); P.t.a = E.t_a.data(); P.t.b = E.t_b.data(); P.t.seq = E.t_seq.data(); P.t.orig = E.t_orig.data();
    P.mode = 1; P.results = out; P.cigar = cigar_buf; P.cigar_cap = cigar_buf ? cigar_cap : 0; P.ncigar = &ncig;
    if (packed && n) { emu_affine_round(E, match, mismatch, gap_open, gap_extend, kt, okt); P.res = E.res.data(); }
    std::vector<uint8_t> dir;
    int bad = 0;
    for (int64_t t = 0; t < n; t++)
        if (emu_ssw_task(P, (int)t, E.qlen[E.t_seq[t]], dir, nullptr) != 1) bad = 1;
    if (bad) return LSW_E_INTERNAL;
    if (cigar_used) *cigar_used = (int64_t)ncig;
    if (cigar_buf && (int64_t)ncig > cigar_cap) return LSW_E_OVERFLOW;
    return LSW_OK;
}


// lsw_chain.cuh on the host: the sub-reads of stage 1 for hits that already carry the survivor flag
int lsw_emu_chain_subreads(const lsw_hit* hits, int64_t n_hits, int64_t n_reads, const int32_t* rlen, int n_fwd, const int32_t* fwd_order,
                           int n_rev, const int32_t* rev_order, int t_query, int tc_query, lsw_subread* out, int64_t cap, int64_t* n_out) {
    ChainOrders C;
    memset(&C, 0, sizeof C);
    C.n_fwd = n_fwd; C.n_rev = n_rev; C.t_query = t_query; C.tc_query = tc_query; C.fwd_t = -1; C.rev_tc = -1;
    for (int i = 0; i < n_fwd; i++) { C.fwd[i] = fwd_order[i]; if (fwd_order[i] == t_query && C.fwd_t < 0) C.fwd_t = i; }
    for (int i = 0; i < n_rev; i++) { C.rev[i] = rev_order[i]; if (rev_order[i] == tc_query && C.rev_tc < 0) C.rev_tc = i; }
    if (C.fwd_t < 0 || C.rev_tc < 0) return LSW_E_ARG;
    int64_t total = 0, lo = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        while (lo < n_hits && hits[lo].read < r) lo++;
        int64_t hi = lo;
        while (hi < n_hits && hits[hi].read == r) hi++;
        const int k = chain_subreads_read(hits, lo, hi, rlen[r], C, nullptr);
        if (total + k <= cap) chain_subreads_read(hits, lo, hi, rlen[r], C, out + total);
        total += k;
        lo = hi;
    }
    if (n_out) *n_out = total;
    return total > cap ? LSW_E_OVERFLOW : LSW_OK;
}


// stage 2 (lsw_ssw.cuh ssw_generic_task) on the host, one sub-read at a time
int lsw_emu_ssw_subreads(int64_t n_reads, const uint8_t* ascii, const int64_t* offs, int match, int mismatch, int gap_open,
                         int gap_extend, int64_t n, const lsw_pair* pairs, const uint8_t* target, int32_t target_len,
                         lsw_result* out, uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used) {
    Emu E;
    const uint8_t q1[1] = {'A'};
    const int32_t qo[2] = {0, 1};
    if (!setup(E, n_reads, ascii, offs, 1, q1, qo, 2, -1, -1)) return LSW_E_UNSUPPORTED;
    int maxq = 1;
    for (int64_t i = 0; i < n; i++) {
        if (pairs[i].read < 0 || pairs[i].read >= n_reads || pairs[i].start < 0 || pairs[i].end < pairs[i].start ||
            pairs[i].end > E.rlen[pairs[i].read]) return LSW_E_ARG;
        maxq = std::max(maxq, pairs[i].end - pairs[i].start);
    }
    std::vector<uint8_t> tcodes((size_t)std::max(target_len, 1));
    for (int i = 0; i < target_len; i++) tcodes[i] = encode_base(target[i]);
    const long long dir_cap = 256 << 10;
    std::vector<uint32_t> he((size_t)maxq);
    std::vector<uint8_t> qbuf((size_t)maxq), dir((size_t)dir_cap);
    unsigned long long ncig = 0;
    int32_t errflag = 0;
    SswGenParams G;
    memset(&G, 0, sizeof G);
    G.codes = E.codes.data(); G.roff = E.roff.data(); G.pairs = pairs; G.n = n; G.tcodes = tcodes.data(); G.tlen = target_len;
    G.match = match; G.mismatch = mismatch; G.gap_open = gap_open; G.gap_extend = gap_extend;
    G.he = he.data(); G.qbuf = qbuf.data(); G.qcap = maxq; G.dir = dir.data(); G.dir_cap = dir_cap;
    G.results = out; G.cigar = cigar_buf; G.cigar_cap = cigar_buf ? cigar_cap : 0; G.ncigar = &ncig; G.errflag = &errflag;
    for (int64_t k = 0; k < n; k++)
        if (ssw_generic_task(G, k, 0, 1) != 1) errflag = 1;
    if (errflag) return LSW_E_UNSUPPORTED;
    if (cigar_used) *cigar_used = (int64_t)ncig;
    if (cigar_buf && (int64_t)ncig > cigar_cap) return LSW_E_OVERFLOW;
    return LSW_OK;
}

}  // extern "C"
