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


}  // namespace lsw
