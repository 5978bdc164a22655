// lsw_host.hpp -- host-side pieces shared by the product library (lsw.cu) and the CPU emulation
// of the kernels used by the tests (lsw_emu.cu).
#pragma once
#include <stdint.h>
#include <vector>

namespace lsw {

constexpr int NCLASS = 9;
static const int CLASS_NR[NCLASS] = {16, 20, 24, 28, 32, 40, 48, 56, 64};    // register-row classes of K1 and K2-fast
static const int CLASS_NR2[NCLASS] = {16, 32, 32, 32, 32, 48, 48, 64, 64};   // rows of K2-full (one op word per 16 rows)

// cand_ok[s][score]: can a task of sequence s whose best local score is `score` still pass
// findPrimerAlignments' test  identity >= thr and span >= len(primer) * thr  (lamprey.py:616-617)?
// score = match*m + mismatch*mm + gap*(I+D) exactly (every cell on a traceback path satisfies
// H == predecessor + step), identity = m / (m + mm + I + D), span <= m + mm + (I+D).  The table is a
// NECESSARY condition evaluated with the same double arithmetic Python uses, so pruning is exact.
// Entry: 0 = no alignment with that score can pass; otherwise 1 + the LARGEST error count x = mm + I + D of any
// passing alignment with that score (K2 drops a candidate whose fewest possible errors already exceed it, lsw_band.cuh).
// all = true (align_tasks: every task is traced and nothing is dropped): 255.
inline void build_cand_ok(int nseq, const int32_t* qlen, int match, int mismatch, int g, double thr, bool all,
                          std::vector<uint8_t>& tab) {
    tab.assign((size_t)nseq * 128, 0);
    for (int s = 0; s < nseq; s++) {
        const int Q = qlen[s];
        for (int score = 0; score < 128; score++) {
            int best = 0;
            if (all) best = 255;
            else if (score > 0) {
                for (int m = 0; m <= Q; m++)
                    for (int mm = 0; m + mm <= Q; mm++) {
                        const int rest = match * m + mismatch * mm - score;   // = g * gaps
                        if (rest < 0 || rest % g) continue;
                        const int gaps = rest / g;
                        const int x = mm + gaps, tot = m + x;
                        if (tot <= 0) continue;
                        const double identity = (double)m / (double)tot;
                        if (identity >= thr && (double)(m + mm + gaps) >= (double)Q * thr) {
                            const int v = x + 1 > 254 ? 254 : x + 1;
                            if (v > best) best = v;
                        }
                    }
            }
            tab[(size_t)s * 128 + score] = (uint8_t)best;
        }
    }
}

// Head / warm-up length of a child slice per sequence (lsw_reuse.cuh, `Wx`).  A child slice [a, b) may take the whole read's
// values for read columns >= a + Wx if every cell there that MATTERS is the same in both matrices.  Cells that matter are those
// whose value could pass the hit test, H >= minpass = the smallest score with cand_ok != 0: an alignment with score s >= minpass
// has at most (match * Q - s) / g deleted read bases (s <= match * Q - g * D), so it spans at most Q + (match * Q - minpass) / g
// read columns and lies inside the slice once its end is that far from `a` -- the cell then holds the same value, at the same
// place, in both matrices.  Cells below minpass may differ (the whole read's value can be larger, a warm-started tail's smaller),
// but they stay below minpass in both, and a task whose best cell is below minpass yields no hit and no children whatever that
// cell is.  With minpass = 1 this is the a-priori bound Q + match * Q / g (every positive cell exact).
inline void build_reuse_wx(int nseq, const int32_t* qlen, int match, int g, const std::vector<uint8_t>& cand_ok, bool full,
                           std::vector<int32_t>& wx) {
    wx.assign((size_t)nseq, 0);
    for (int s = 0; s < nseq; s++) {
        const int Q = qlen[s], top = match * Q;
        int minpass = top > 0 ? top : 1;                    // nothing can pass: any bound will do
        for (int score = 1; score < 128 && score <= top; score++)
            if (cand_ok[(size_t)s * 128 + score]) { minpass = score; break; }
        if (full) minpass = 1;
        const int apriori = Q + top / g;
        const int w = Q + (top - minpass) / g + 1;
        wx[(size_t)s] = w < apriori ? w : apriori;
    }
}

}  // namespace lsw
