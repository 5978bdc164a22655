// lsw_affine.cuh -- K1 for the DEFAULT seeding path: the forward scan of ssw.c's ssw_align (scikit-bio StripedSmithWaterman,
// /root/reference/lamprey.py:572-593) with packed s16x2 DPX instructions.
//
// Same machinery as ScoreLane (lsw_score.cuh): one lane owns two tasks, one per s16 half; all primer rows of both live in
// registers; the lane streams read columns, retires finished tasks and claims new ones every 64 columns; the warp's profile
// sits in shared memory.  What differs is the recurrence and the arg-max rule:
//
//   affine gaps       H = max(0, D + w, E, F),  E' = max(E - gapE, H - gapO),  F' = max(F - gapE, H - gapO)  (all floored at 0,
//                     as ssw.c's unsigned saturating arithmetic does).  Values are scaled by 256 (the low byte of a half
//                     carries the arg-max tag) but NOT biased: every packed add is a true two-lane DPX add, so a negative
//                     substitution score cannot borrow across the halves.  Per two cells:
//                         t  = VIADDMNMX.RELU(D, w, E)      h  = VIMNMX(t, F)
//                         hg = VIADDMNMX.RELU(h, -gapO, -gapO)  E' = VIADDMNMX(E, -gapE, hg)   F' = VIADDMNMX(F, -gapE, hg)
//   first maximum     ssw.c moves its best score only when a column's maximum is STRICTLY greater than everything before it,
//                     and takes the SMALLEST read index of that column: the arg-max is the lexicographic maximum of
//                     (H, -column, -row).  The tag in the low byte is 2 * (63 - step) + (1 - row parity), the fold orders row
//                     pairs by descending row, and a later block replaces the best only with a strictly greater value.
//   "N"               a non-ACGT read base scores 0 against everything (skbio's _build_match_matrix); padding rows beyond the
//                     primer score -64 so that they can never hold a column's maximum.
// Requires match * len(primer) <= 127 (8.8 fixed point in an s16 half), as the swalign path does.
#pragma once
#include "lsw_common.cuh"
#include "lsw_score.cuh"

namespace lsw {

LSW_HD uint32_t addmax2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }   // max(a + b, c, 0)
LSW_HD uint32_t max2_2(uint32_t a, uint32_t b) {                                                               // max per s16 half
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);                 // VIMNMX.S16x2 with two operands: twice the rate of the three-operand forms
#else
    return __vimax3_s16x2(a, b, b);
#endif
}

template <int NR>
struct AffineLane {
    static_assert(NR % 4 == 0 && NR <= 64, "rows are kept in registers");
    static constexpr int PROF_STRIDE = NR / 2 + 1;
    static constexpr int PAD_SCORE = -64;      // substitution score of a padding row (rows >= len(primer))

    uint32_t M[NR];        // 256 * H of the previous column, two tasks packed
    uint32_t E[NR];        // 256 * E for the current column
    uint32_t K[NR / 2];    // per row pair: running max of (256 * H + tag) over the current block
    int32_t  tid[2];
    int64_t  pos[2];
    int32_t  begin[2];
    int32_t  rem[2];
    int32_t  cbase[2];
    uint32_t best[2];      // (H << 14) | (63 - step) << 6 | (63 - row) of the best cell so far; 0 = nothing positive yet
    int32_t  bestc[2];     // its DP column (1-based)
    int32_t  len[2];
    U4       nx[2];
    unsigned long long n_tasks, n_cells, n_steps, n_comp;

    LSW_HD void init() {
#pragma unroll
        for (int r = 0; r < NR; r++) { M[r] = 0; E[r] = 0; }
#pragma unroll
        for (int r = 0; r < NR / 2; r++) K[r] = 0;
        tid[0] = tid[1] = -1;
        rem[0] = rem[1] = 0; begin[0] = begin[1] = 0;
        pos[0] = pos[1] = 0; cbase[0] = cbase[1] = 0;
        best[0] = best[1] = 0; bestc[0] = bestc[1] = 0; len[0] = len[1] = 0;
        nx[0].x = nx[0].y = nx[0].z = nx[0].w = 0;
        nx[1] = nx[0];
    }

    // packed profile of sequence q: row idx (NIDX layout of lsw_common.cuh) holds (w(code_lo, q[r]), w(code_hi, q[r])) * 256.
    // P.s_match / P.s_mis carry 256 * match / 256 * mismatch here (no gap bias).
    static LSW_HD void build_profile(const ScoreParams& P, int q, int Q, uint32_t* prof32, int lane, int nlanes) {
        const uint8_t* qc = P.qcodes + (size_t)q * QSTRIDE;
        for (int e = lane; e < NIDX * NR; e += nlanes) {
            const int idx = e / NR, r = e - idx * NR;
            int cl, ch;
            if (idx < 16) { cl = idx >> 2; ch = idx & 3; }
            else if (idx < 20) { cl = 4; ch = idx - 16; }
            else if (idx == 36) { cl = 4; ch = 4; }
            else { cl = (idx - 20) >> 2; ch = 4; }
            int lo, hi;
            if (r >= Q) { lo = hi = PAD_SCORE * SC; }
            else {
                const int qb = qc[r];
                lo = cl == 4 ? 0 : (cl == qb ? P.s_match : P.s_mis);
                hi = ch == 4 ? 0 : (ch == qb ? P.s_match : P.s_mis);
            }
            prof32[idx * (PROF_STRIDE * 2) + r] = ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
        }
    }

    LSW_HD void start(const ScoreParams& P, int h, int t, const KTask& k, int Q) {
        const int64_t p0 = k.p0;
        tid[h] = t;
        len[h] = k.len;
        rem[h] = k.len;
        begin[h] = (int32_t)(p0 & 15);
        pos[h] = p0 & ~(int64_t)15;
        cbase[h] = 1 - begin[h];
        best[h] = 0;
        bestc[h] = 0;
        const uint32_t keep = h ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
        for (int r = 0; r < NR; r++) { M[r] &= keep; E[r] &= keep; }
        nx[h] = *reinterpret_cast<const U4*>(P.codes + pos[h]);
        if (P.count_ref) {
            n_tasks += 1;
            n_cells += (unsigned long long)k.len * (unsigned long long)Q;
        }
        n_comp += (unsigned long long)k.len * (unsigned long long)Q;
    }

    // publish (score, end row + 1, end column + 1) -- 0 / 0 / 0 if nothing aligned -- and queue the task for the traceback
    template <class PushCand>
    LSW_HD void finish(const ScoreParams& P, int h, int q, int Q, PushCand push) {
        const int t = tid[h];
        int score = 0, row = 0, col = 0;
        if (best[h]) {
            score = (int)(best[h] >> 14);
            row = 63 - (int)(best[h] & 63u) + 1;
            col = bestc[h];
        }
        P.res[t] = make_int2(score | (row << 16), col);
        if (P.push_cands && P.cand_ok[(size_t)q * 128 + score]) push(t, 0, score | (row << 16), col);
        tid[h] = -1;
    }

    // two DP columns for both halves.  tagA / tagB: 2 * (63 - step) in both halves for the two columns.
    LSW_HD void column2(const U2* prowA, const U2* prowB, uint32_t tagA, uint32_t tagB, uint32_t NGO2, uint32_t NGE2) {
        uint32_t dOld = 0, uA = 0, fA = 0, fB = 0;      // dOld: M of row r-1 in the column before A; uA: h of row r-1 in column A
#pragma unroll
        for (int r2 = 0; r2 < NR / 2; r2++) {
            const U2 sa = prowA[r2], sb = prowB[r2];
#define LSW_AROW2(R, SA, SB)                                              \
            {                                                             \
                const uint32_t mold = M[R];                               \
                const uint32_t h1 = max2_2(addmax2_relu(dOld, SA, E[R]), fA);      /* column A */ \
                const uint32_t g1 = addmax2_relu(h1, NGO2, NGO2);      /* max(h - go, -go, 0): a zero operand would be re-made (PRMT) every time */           \
                const uint32_t e1 = addmax2(E[R], NGE2, g1);              \
                fA = addmax2(fA, NGE2, g1);                               \
                const uint32_t h2 = max2_2(addmax2_relu(uA, SB, e1), fB);          /* column B */ \
                const uint32_t g2 = addmax2_relu(h2, NGO2, NGO2);           \
                E[R] = addmax2(e1, NGE2, g2);                             \
                fB = addmax2(fB, NGE2, g2);                               \
                K[(R) >> 1] = max3_2(K[(R) >> 1], h1 + (tagA + (1 - ((R) & 1)) * 0x00010001u), \
                                     h2 + (tagB + (1 - ((R) & 1)) * 0x00010001u)); \
                M[R] = h2;                                                \
                dOld = mold; uA = h1;                                     \
            }
            LSW_AROW2(r2 * 2 + 0, sa.x, sb.x)
            LSW_AROW2(r2 * 2 + 1, sa.y, sb.y)
#undef LSW_AROW2
        }
    }

    LSW_HD void run_block(const ScoreParams& P, const U2* prof, const U4* keep_ge, const U4* keep_lt, int Q, int q) {
        (void)q;
        const uint32_t NGO2 = ((uint32_t)(-(P.aff_open * SC)) & 0xFFFFu) * 0x00010001u;
        const uint32_t NGE2 = ((uint32_t)(-(P.aff_extend * SC)) & 0xFFFFu) * 0x00010001u;
        int end[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            end[h] = 0;
            if (tid[h] >= 0) { end[h] = begin[h] + rem[h]; if (end[h] > BLOCK_STEPS) end[h] = BLOCK_STEPS; }
            else begin[h] = 0;
        }
#pragma unroll
        for (int r = 0; r < NR / 2; r++) K[r] = 0;
        uint32_t tag = 2u * 63u * 0x00010001u;            // 2 * (63 - step), both halves
#pragma unroll 1
        for (int k = 0; k < BLOCK_STEPS / 16; k++) {
            U4 c[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                U4 v = nx[h];
                if (tid[h] >= 0) nx[h] = *reinterpret_cast<const U4*>(P.codes + pos[h] + 16 * (k + 1));
                int nl = begin[h] - 16 * k; nl = nl < 0 ? 0 : (nl > 16 ? 16 : nl);
                int ne = end[h] - 16 * k;   ne = ne < 0 ? 0 : (ne > 16 ? 16 : ne);
                const U4 kg = keep_ge[nl], kl = keep_lt[ne];
                const uint32_t OTH = 0x01010101u * (h ? CODE_OTHER_HI : CODE_OTHER);
                if (h) {
                    v.x |= (v.x & 0x04040404u) << 2; v.y |= (v.y & 0x04040404u) << 2;
                    v.z |= (v.z & 0x04040404u) << 2; v.w |= (v.w & 0x04040404u) << 2;
                }
                uint32_t kp;
                kp = kg.x & kl.x; v.x = (v.x & kp) | (OTH & ~kp);
                kp = kg.y & kl.y; v.y = (v.y & kp) | (OTH & ~kp);
                kp = kg.z & kl.z; v.z = (v.z & kp) | (OTH & ~kp);
                kp = kg.w & kl.w; v.w = (v.w & kp) | (OTH & ~kp);
                c[h] = v;
            }
#pragma unroll 1
            for (int w = 0; w < 4; w++) {
                const uint32_t idx4 = c[0].x * 4u + c[1].x;
                c[0].x = c[0].y; c[0].y = c[0].z; c[0].z = c[0].w;
                c[1].x = c[1].y; c[1].y = c[1].z; c[1].z = c[1].w;
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const uint32_t ia = (idx4 >> (8 * j)) & 0xFFu, ib = (idx4 >> (8 * j + 8)) & 0xFFu;
                    column2(prof + ia * PROF_STRIDE, prof + ib * PROF_STRIDE, tag, tag - 0x00020002u, NGO2, NGE2);
                    tag -= 0x00040004u;
                }
            }
        }
        // fold: per half the lexicographic max of (H, 63 - step, 63 - row) over the rows of the primer
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t bb = 0;
#pragma unroll
            for (int r2 = 0; r2 < NR / 2; r2++) {
                if (2 * r2 < Q) {
                    const uint32_t kv = h ? (K[r2] >> 16) : (K[r2] & 0xFFFFu);       // H << 8 | 2 * (63 - step) + (1 - parity)
                    const uint32_t Hh = kv >> 8, L = kv & 0xFFu;
                    const uint32_t key = (Hh << 14) | ((L >> 1) << 6) | ((uint32_t)(31 - r2) << 1) | (L & 1u);
                    bb = key > bb ? key : bb;
                }
            }
            if (tid[h] >= 0 && (bb >> 14) > (best[h] >> 14)) {        // strictly greater: an earlier block keeps ties
                best[h] = bb;
                bestc[h] = cbase[h] + 63 - (int32_t)((bb >> 6) & 63u);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (tid[h] >= 0) {
                rem[h] -= end[h] - begin[h];
                begin[h] = 0;
                pos[h] += BLOCK_STEPS;
                cbase[h] += BLOCK_STEPS;
            }
        }
        n_steps += BLOCK_STEPS;
    }
};

#if defined(__CUDACC__)
template <int NR> struct AffineOcc { static constexpr int value = NR <= 20 ? 3 : NR <= 32 ? 2 : 1; };

template <int NR>
__global__ void __launch_bounds__(SCORE_WARPS * 32, AffineOcc<NR>::value)
k_score_affine(const ScoreParams P) {
    score_kernel_body<NR, AffineLane<NR>>(P);
}
#endif

}  // namespace lsw
