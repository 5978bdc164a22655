// lsw_score.cuh -- K1: score + arg-max pass of swalign.LocalAlignment.align()
//
// Replaces the matrix fill and the `if val >= max_val` arg-max of the third-party align()
// that /root/reference/lamprey.py:563 calls (semantics: SURVEY.md Appendix A.1 / A.3).
// For LAMPrey's parameters gap_penalty == gap_extension_penalty, so the VALUE recurrence is
// linear-gap Smith-Waterman,  H = max(0, D + s, L - g, U - g);  the (op, run) part of
// swalign's cells only matters for the traceback (K2, lsw_trace.cuh).
//
// Layout of the computation (see DESIGN.md section "K1"):
//  * one LANE owns two independent tasks at a time, one in each s16 half of every register
//    (packed DPX: VIADDMNMX.S16x2 / VIMNMX3.S16x2).  All NR query rows of both tasks live in
//    registers; the lane streams read columns.
//  * all lanes of a WARP serve the same schema sequence, whose substitution profile
//    (indexed by the pair of read codes of the two tasks) sits in shared memory.
//  * values are scaled by 256 and biased by |gap| so that
//        T     = max(Mold[r-1] + s'', Mold[r])              s'' = 256*(s + g)
//        h     = max(T, Mnew[r-1], 256*g)                   h   = 256*(H + g)
//        Mnew  = h - 256*g                                  (= 256*H, never borrows)
//        K[r/2] = max(K[r/2], h + 64*(r&1) + step)          arg-max key of a row pair: row bit and step in the low byte
//    i.e. 3 ALU-pipe DPX ops + 1 add per TWO cells.
//  * every BLOCK_STEPS columns the lane folds K into its per-task best (value, row, column)
//    with swalign's tie rule (last cell in row-major order wins: highest row, then highest
//    column), retires finished tasks and claims new ones from the per-sequence queue.
#pragma once
#include "lsw_common.cuh"

namespace lsw {

#ifndef LSW_K1_KEYV
#define LSW_K1_KEYV 0        // how K1 folds the arg-max keys: see ScoreLane::column2
#endif

// a * one + b as a real multiply-add (FMA pipe); `one` is a run-time 1, so that the assembler cannot turn it back into an add
LSW_HD uint32_t mad_one(uint32_t a, uint32_t one, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t o;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(o) : "r"(a), "r"(one), "r"(b));
    return o;
#else
    return a * one + b;
#endif
}

// a * 0x10001 + b as a multiply-add with an IMMEDIATE multiplier: the assembler cannot turn it into an add, so it is issued on
// the FMA pipe (IMAD), next to the DPX instructions that saturate the ALU pipe.  a = a small count: a * 0x10001 puts it into
// both s16 halves; a = -256 g: a * 0x10001 == -(0x01000100 g) (mod 2^32), the packed subtraction of the gap penalty.
LSW_HD uint32_t mad_x2(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t o;
    asm("mad.lo.u32 %0, %1, 0x10001, %2;" : "=r"(o) : "r"(a), "r"(b));
    return o;
#else
    return a * 0x10001u + b;
#endif
}

struct U4 { uint32_t x, y, z, w; };
struct U2 { uint32_t x, y; };

template <int NR>
struct ScoreLane {
    static_assert(NR % 4 == 0 && NR <= 64, "rows are kept in registers, 4 per profile load");
    static constexpr int PROF_STRIDE = NR / 2 + 1;      // 8-byte words per profile row: odd, see NIDX

    uint32_t M[NR];        // 256 * H of the previous column, two tasks packed
    uint32_t K[NR / 2];    // per ROW PAIR: running max of (256*(H+g) + 64*(row & 1) + step) over the current block
    int32_t  tid[2];       // task index, -1 = half idle
    int64_t  pos[2];       // byte offset (16-aligned) in codes of step 0 of the next block
    int32_t  begin[2];     // first live step of the next block (only the first block of a task is > 0)
    int32_t  rem[2];       // live columns still to run
    int32_t  cbase[2];     // DP column (1-based) of step 0 of the next block
    uint32_t best[2];      // ((H+g) << 16) | (row << 8) | step of the best cell so far
    int32_t  bestc[2];     // its DP column
    int32_t  len[2];       // slice length
    int32_t  skipb[2];     // leading blocks of the current task that are warm-up only (not folded)
    uint32_t bbo[2];       // round 0: first block-best slot of the task's read (lsw_reuse.cuh), fetched when the task starts
    U4       nx[2];        // read codes of the next 16 steps, fetched one chunk ahead
    unsigned long long n_tasks, n_cells, n_steps, n_comp;

    LSW_HD void init() {
#pragma unroll
        for (int r = 0; r < NR; r++) M[r] = 0;
#pragma unroll
        for (int r = 0; r < NR / 2; r++) K[r] = 0;
        tid[0] = tid[1] = -1;
        rem[0] = rem[1] = 0; begin[0] = begin[1] = 0;
        pos[0] = pos[1] = 0; cbase[0] = cbase[1] = 0;
        best[0] = best[1] = 0; bestc[0] = bestc[1] = 0; len[0] = len[1] = 0; skipb[0] = skipb[1] = 0; bbo[0] = bbo[1] = 0;
        nx[0].x = nx[0].y = nx[0].z = nx[0].w = 0;
        nx[1] = nx[0];
    }

    // Build the packed profile of sequence q for this warp: row idx (see NIDX) holds, for every query
    // row r, (s''(code_lo, q[r]), s''(code_hi, q[r])).  `lane`/`nlanes` split the work.
    static LSW_HD void build_profile(const ScoreParams& P, int q, int Q, uint32_t* prof32, int lane, int nlanes) {
        const uint8_t* qc = P.qcodes + (size_t)q * QSTRIDE;
        for (int e = lane; e < NIDX * NR; e += nlanes) {
            const int idx = e / NR, r = e - idx * NR;
            int cl, ch;                                   // 0..3, or 4 = other
            if (idx < 16) { cl = idx >> 2; ch = idx & 3; }
            else if (idx < 20) { cl = 4; ch = idx - 16; }
            else if (idx == 36) { cl = 4; ch = 4; }
            else { cl = (idx - 20) >> 2; ch = 4; }        // 20, 24, 28, 32 (other values are never indexed)
            int lo = P.s_mis, hi = P.s_mis;
            if (r < Q && cl < 4) {
                const int qb = qc[r];
                if (cl == qb) lo = P.s_match;
                if (ch == qb) hi = P.s_match;
            } else if (r < Q) {
                if (ch == qc[r]) hi = P.s_match;
            }
            prof32[idx * (PROF_STRIDE * 2) + r] = ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16);
        }
    }

    // Start task `t` (descriptor k) in half h.
    LSW_HD void start(const ScoreParams& P, int h, int t, const KTask& k, int Q) {
        const int64_t p0 = k.p0;
        tid[h] = t;
        len[h] = k.len;
        rem[h] = k.len;
        begin[h] = (int32_t)(p0 & 15);
        pos[h] = p0 & ~(int64_t)15;
        cbase[h] = 1 - begin[h];
        best[h] = ((uint32_t)P.g << 16) | 0xFFFFu;     // nothing positive seen yet
        bestc[h] = 0;
        const uint32_t keep = h ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
        for (int r = 0; r < NR; r++) M[r] &= keep;     // zero boundary column for this half
        nx[h] = *reinterpret_cast<const U4*>(P.codes + pos[h]);
        skipb[h] = k.skip;
        // the row of block bests this task writes: looked up HERE, a block before the first store needs it (the address used to be
        // chased through task -> read -> offset at every store: 12 % of round 0 sat on those dependent loads)
        if (P.blockbest) bbo[h] = (uint32_t)P.bb_off[P.t.read[t]];
        if (P.count_ref) {
            n_tasks += 1;
            n_cells += (unsigned long long)k.len * (unsigned long long)Q;
        }
        n_comp += (unsigned long long)k.len * (unsigned long long)Q;
    }

    // Retire the task in half h: publish (score, max_row, max_col) and queue it for traceback.
    template <class PushCand>
    LSW_HD void finish(const ScoreParams& P, int h, int q, int Q, PushCand push) {
        const int t = tid[h];
        int score = (int)(best[h] >> 16) - P.g;
        int row, col;
        if (bestc[h] == 0) {            // no positive cell: swalign's `>=` leaves the last cell
            score = 0;
            row = len[h] > 0 ? Q : 0;
            col = len[h] > 0 ? len[h] : 0;
        } else {
            row = (int)((best[h] >> 8) & 0xFF) + (int)((best[h] >> 6) & 1) + 1;
            col = bestc[h];
        }
        P.res[t] = make_int2(score | (row << 16), col);
        if (P.push_cands && P.cand_ok[(size_t)q * 128 + score]) {
            // sort key for K2: the number of window columns its traceback will fill (similar work per warp)
            int w = score > 0 ? row + (P.match * row - score) / P.g + 8 : 0;
            if (w > col) w = col;
            push(t, w, score | (row << 16), col);
        }
        tid[h] = -1;
    }

    // TWO DP columns (c, c+1) for both halves, row by row.  prowA / prowB: profile rows of the two columns'
    // (code_lo, code_hi) pairs.  Per row and column pair (four cells): 2 VIADDMNMX + 2 VIMNMX3 for the values, two subtractions,
    // and the arg-max keys of both columns folded into K[r].  KEYV selects how the keys are folded (all variants compute the
    // same K; they differ in which pipe the instructions land on -- measured by lsw_measure_int_peak modes 9..12):
    //   0  K = max3(K, h1 + tagA, h2 + tagB)            2 adds + 1 VIMNMX3   (the compiler issues the adds as VIADD: ALU pipe)
    //   1  K = max(h2 + tagB, max(h1 + tagA, K))        2 VIADDMNMX          (one ALU instruction less, none on the FMA pipe)
    //   2  as 0 with the adds forced onto the FMA pipe  (IMAD with a run-time multiplier of one)
    //   3  as 0 with the key adds as IMAD with the immediate multiplier 0x10001 (mad_x2): FMA pipe
    //   4  as 3, and the two subtractions of the gap penalty the same way
    template <int KEYV = LSW_K1_KEYV>
    LSW_HD void column2(const U2* prowA, const U2* prowB, uint32_t tvA, uint32_t tvB, uint32_t G2, uint32_t FLOOR2, uint32_t one = 1u) {
        uint32_t dOld = 0, uA = 0, uB = 0;
        const uint32_t negG = 0u - (G2 & 0xFFFFu) * one;       // -256 g (times a run-time one: stays a register operand)
#pragma unroll
        for (int r2 = 0; r2 < NR / 2; r2++) {
            const U2 sa = prowA[r2], sb = prowB[r2];
#define LSW_ROW2(R, SA, SB)                                      \
            {                                                    \
                const uint32_t mold = M[R];                      \
                const uint32_t T1 = addmax2(dOld, SA, mold);     \
                const uint32_t h1 = max3_2(T1, uA, FLOOR2);      \
                const uint32_t m1 = KEYV == 4 ? mad_x2(negG, h1) : h1 - G2; \
                const uint32_t T2 = addmax2(uA, SB, m1);         \
                const uint32_t h2 = max3_2(T2, uB, FLOOR2);      \
                const uint32_t m2 = KEYV == 4 ? mad_x2(negG, h2) : h2 - G2; \
                const uint32_t tA = tvA + ((R) & 1) * 0x00400040u, tB = tvB + ((R) & 1) * 0x00400040u; \
                if (KEYV == 3 || KEYV == 4) K[(R) >> 1] = max3_2(K[(R) >> 1], mad_x2((tvA & 0xFFFFu) + ((R) & 1) * 64u, h1), mad_x2((tvB & 0xFFFFu) + ((R) & 1) * 64u, h2)); \
                else if (KEYV == 1) K[(R) >> 1] = addmax2(h2, tB, addmax2(h1, tA, K[(R) >> 1])); \
                else if (KEYV == 2) K[(R) >> 1] = max3_2(K[(R) >> 1], mad_one(h1, one, tA), mad_one(h2, one, tB)); \
                else K[(R) >> 1] = max3_2(K[(R) >> 1], h1 + tA, h2 + tB); \
                M[R] = m2;                                       \
                dOld = mold; uA = m1; uB = m2;                   \
            }
            LSW_ROW2(r2 * 2 + 0, sa.x, sb.x)
            LSW_ROW2(r2 * 2 + 1, sa.y, sb.y)
#undef LSW_ROW2
        }
    }

    // P.block_steps (64, or 32 for the short segment tasks of lsw_reuse.cuh) columns for both halves, then fold the block's arg-max into the task state.
    // keep_ge[n]: bytes j >= n of a 16-byte chunk are live; keep_lt[n]: bytes j < n are live.
    LSW_HD void run_block(const ScoreParams& P, const U2* prof, const U4* keep_ge, const U4* keep_lt, int Q, int q) {
        const uint32_t G2 = (uint32_t)(P.g * SC) * 0x00010001u;
        const uint32_t FLOOR2 = G2;
        int end[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            end[h] = 0;
            if (tid[h] >= 0) { end[h] = begin[h] + rem[h]; if (end[h] > P.block_steps) end[h] = P.block_steps; }
            else begin[h] = 0;
        }
        // Round 0 (P.blockbest set; whole reads, 64-step blocks) remembers the best cell of every BB_BLOCK = 32 columns for the
        // child tasks (lsw_reuse.cuh): the block is folded twice, after each half.  Otherwise one fold per block.
        const int nparts = P.blockbest ? 2 : 1;
        const int kper = P.block_steps / (16 * nparts);
        uint32_t tvec = 0;
#pragma unroll 1
        for (int part = 0; part < nparts; part++) {
#pragma unroll
            for (int r = 0; r < NR / 2; r++) K[r] = 0;
#pragma unroll 1
            for (int k = part * kper; k < (part + 1) * kper; k++) {
                U4 c[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    U4 v = nx[h];
                    // fetch the next chunk now (the first chunk of the next block when k == 3: the buffer is
                    // padded, and start() reloads if the half switches task); its latency hides behind 16 columns
                    if (tid[h] >= 0) nx[h] = *reinterpret_cast<const U4*>(P.codes + pos[h] + 16 * (k + 1));
                    int nl = begin[h] - 16 * k; nl = nl < 0 ? 0 : (nl > 16 ? 16 : nl);
                    int ne = end[h] - 16 * k;   ne = ne < 0 ? 0 : (ne > 16 ? 16 : ne);
                    const U4 kg = keep_ge[nl], kl = keep_lt[ne];
                    const uint32_t OTH = 0x01010101u * (h ? CODE_OTHER_HI : CODE_OTHER);
                    if (h) {      // a real non-ACGT base is stored as 4; the high half indexes it as 20
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
                    const uint32_t idx4 = c[0].x * 4u + c[1].x;          // four profile row indices, bytes <= 36
                    c[0].x = c[0].y; c[0].y = c[0].z; c[0].z = c[0].w;
                    c[1].x = c[1].y; c[1].y = c[1].z; c[1].z = c[1].w;
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        const uint32_t ia = (idx4 >> (8 * j)) & 0xFFu, ib = (idx4 >> (8 * j + 8)) & 0xFFu;
                        column2(prof + ia * PROF_STRIDE, prof + ib * PROF_STRIDE, tvec, tvec + 0x00010001u, G2, FLOOR2);
                        tvec += 0x00020002u;
                    }
                }
            }

            // fold: per half, lexicographic max of (H, row, step) over the rows of the query
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t bb = 0;
#pragma unroll
                for (int r2 = 0; r2 < NR / 2; r2++) {
                    if (2 * r2 < Q) {
                        // bytes of K[r2]: [low_lo, H_lo, low_hi, H_hi] with low = 64 * (row & 1) + step
                        //   ->  [low, 2 * r2, H, 0]: ordered by (H, row, step) because 256 * (2 r2) + low is monotone in (row, step).
                        // (If Q is odd the padding row Q can be the pair's best only with a value below a real cell's: see DESIGN.md.)
                        const uint32_t key = prmt(K[r2], (uint32_t)(2 * r2), h ? 0x5342u : 0x5140u);
                        bb = key > bb ? key : bb;
                    }
                }
                if (tid[h] >= 0 && P.blockbest) {
                    // first column of this half block, 0-based on the read; a half block that starts past the read's end is not stored
                    const int col0 = cbase[h] - 1 + part * BB_BLOCK;
                    if (col0 < len[h])          // (every lane of the warp serves sequence q)
                        P.blockbest[(size_t)q * P.bb_total + bbo[h] + (uint32_t)(col0 / BB_BLOCK)] = bb;
                }
                if (tid[h] >= 0 && skipb[h] == 0 && (bb >> 6) >= (best[h] >> 6)) {   // >= : a later block wins ties on (H, row); bits 6.. hold (H, row)
                    best[h] = bb;                                                    // (warm-up blocks of a tail segment are not counted)
                    bestc[h] = cbase[h] + (int32_t)(bb & 0x3F);
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (tid[h] >= 0) {
                if (skipb[h] > 0) skipb[h]--;
                rem[h] -= end[h] - begin[h];
                begin[h] = 0;
                pos[h] += P.block_steps;
                cbase[h] += P.block_steps;
            }
        }
        n_steps += P.block_steps;
    }
};

// Live-byte masks of a 16-byte chunk, indexed by n in [0, 16].
LSW_HD void make_keep_masks(U4* keep_ge, U4* keep_lt, int n) {
    uint32_t ge[4], lt[4];
    for (int w = 0; w < 4; w++) {
        ge[w] = 0; lt[w] = 0;
        for (int j = 0; j < 4; j++) {
            const int byte = 4 * w + j;
            if (byte >= n) ge[w] |= 0xFFu << (8 * j);
            if (byte < n)  lt[w] |= 0xFFu << (8 * j);
        }
    }
    keep_ge[n].x = ge[0]; keep_ge[n].y = ge[1]; keep_ge[n].z = ge[2]; keep_ge[n].w = ge[3];
    keep_lt[n].x = lt[0]; keep_lt[n].y = lt[1]; keep_lt[n].z = lt[2]; keep_lt[n].w = lt[3];
}

#if defined(__CUDACC__)
// ---- the kernel --------------------------------------------------------------------------
// One launch serves every sequence of one register-row class.  Warps pick the sequence with
// the most unclaimed tasks, build its profile in shared memory, and their lanes drain that
// sequence's queue two tasks at a time until it is empty.
constexpr int SCORE_WARPS = 4;

// resident blocks per SM the register allocator must allow for (65536 / (128 threads * regs))
#ifndef LSW_OCC_BOOST
#define LSW_OCC_BOOST 2      // measured on B200: 128 registers (4 blocks/SM) for NR <= 28 is spill-free and 1.5 % faster; forcing 5-6 blocks spills and costs 20 %
#endif
template <int NR> struct ScoreOcc {
    static constexpr int value = LSW_OCC_BOOST == 1 ? (NR <= 16 ? 6 : NR <= 24 ? 5 : NR <= 32 ? 4 : NR <= 40 ? 3 : 2)
                               : LSW_OCC_BOOST == 2 ? (NR <= 28 ? 4 : NR <= 40 ? 3 : 2)
                               : LSW_OCC_BOOST == 3 ? (NR <= 20 ? 5 : NR <= 28 ? 4 : NR <= 40 ? 3 : 2) : 1;
};

template <int NR, class Lane>
__device__ __forceinline__ void score_kernel_body(const ScoreParams& P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    U4* keep_ge = reinterpret_cast<U4*>(smem_raw);
    U4* keep_lt = keep_ge + 17;
    U2* prof_all = reinterpret_cast<U2*>(keep_lt + 17);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    U2* prof = prof_all + (size_t)warp * NIDX * Lane::PROF_STRIDE;

    if (threadIdx.x < 17) make_keep_masks(keep_ge, keep_lt, threadIdx.x);
    __syncthreads();

    Lane L;
    L.n_tasks = L.n_cells = L.n_steps = L.n_comp = 0;

    // Warps spread over the sequences of the class (warp w starts on sequence w mod n) and stay on a
    // sequence until its queue is empty; only then do they move to the fullest remaining queue.
    // (If every warp chased the fullest queue they would all drain the same one together and
    // all pay the end-of-queue idle time once per sequence instead of once or twice per launch.)
    int prefer = (int)((blockIdx.x * SCORE_WARPS + warp) % (unsigned)P.n_class_seqs);
    for (;;) {
        long long bestrem = 0; int bestq = -1;
        {
            const int q = P.class_seqs[prefer];
            const long long cnt = P.seg[q + 1] - P.seg[q];
            const long long hd = (long long)*((volatile unsigned long long*)&P.head[q]);
            if (hd < cnt) { bestrem = cnt - hd; bestq = q; }
        }
        if (bestq < 0) {
            int besti = -1;
            for (int i = lane; i < P.n_class_seqs; i += 32) {
                const int q = P.class_seqs[i];
                const long long cnt = P.seg[q + 1] - P.seg[q];
                const long long hd = (long long)*((volatile unsigned long long*)&P.head[q]);
                const long long remq = cnt - (hd < cnt ? hd : cnt);
                if (remq > bestrem) { bestrem = remq; bestq = q; besti = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const long long orem = __shfl_xor_sync(0xFFFFFFFFu, bestrem, o);
                const int oq = __shfl_xor_sync(0xFFFFFFFFu, bestq, o);
                const int oi = __shfl_xor_sync(0xFFFFFFFFu, besti, o);
                if (orem > bestrem || (orem == bestrem && oq > bestq)) { bestrem = orem; bestq = oq; besti = oi; }
            }
            if (besti >= 0) prefer = besti;
        }
        if (bestq < 0) break;
        const int q = bestq;
        const int Q = P.qlen[q];
        const long long cnt = P.seg[q + 1] - P.seg[q];
        const long long base = P.seg[q];

        __syncwarp();
        Lane::build_profile(P, q, Q, reinterpret_cast<uint32_t*>(prof), lane, 32);
        __syncwarp();

        L.init();
        bool exhausted = false;
        // the NEXT task of each half is claimed, and its descriptor load issued, one block before it is needed:
        // the load's latency hides behind that block instead of stalling the warp at the task switch
        int nxt[2] = {-1, -1};
        KTask nk[2];
        nk[0].p0 = nk[1].p0 = 0; nk[0].len = nk[1].len = 0; nk[0].skip = nk[1].skip = 0;
        for (;;) {
            // 1. retire finished tasks, activate prefetched ones
            bool want[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (L.tid[h] >= 0 && L.rem[h] == 0)
                    L.finish(P, h, q, Q, [&](int t, int key, int res_x, int col) {
                        const KTask k = P.kt[t];                       // plain task list: p0 = slice start, len = slice length
                        CandRec c;
                        c.p0 = k.p0; c.len = k.len; c.rd = P.t.read[t]; c.a = P.t.a[t];
                        c.res_x = res_x; c.col = col; c.t = t;
                        const unsigned long long i = append_slot(P.ncand);
                        P.cand[i] = (uint32_t)i;
                        P.cand_key[i] = (uint16_t)(((uint32_t)q << 8) | (uint32_t)key);
                        P.cand_rec[i] = c;
                    });
                if (L.tid[h] < 0 && nxt[h] >= 0) { L.start(P, h, nxt[h], nk[h], Q); nxt[h] = -1; }
                // a task is wanted now (idle half) or for the next boundary (this block is the current task's last)
                want[h] = !exhausted && nxt[h] < 0 &&
                          (L.tid[h] < 0 || L.rem[h] <= P.block_steps - L.begin[h]);
            }
            // 2. claim tasks for the whole warp with ONE atomic: all lanes pull from the same per-sequence queue, and
            //    per-lane atomics on one address serialise in L2 (tens of millions of short tasks per round)
            const unsigned b0 = __ballot_sync(0xFFFFFFFFu, want[0]), b1 = __ballot_sync(0xFFFFFFFFu, want[1]);
            if (b0 | b1) {
                const int c0 = __popc(b0), c1 = __popc(b1);
                unsigned long long first = 0;
                if (lane == 0) first = atomicAdd(&P.head[q], (unsigned long long)(c0 + c1));
                first = __shfl_sync(0xFFFFFFFFu, first, 0);
                const unsigned lt = (1u << lane) - 1u;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (!want[h]) continue;
                    const long long i = (long long)first + (h ? c0 + __popc(b1 & lt) : __popc(b0 & lt));
                    if (i >= cnt) { exhausted = true; continue; }
                    const int t = (int)(base + i);
                    if (L.tid[h] < 0) L.start(P, h, t, P.kt[t], Q);           // idle half: start right away
                    else { nxt[h] = t; nk[h] = P.kt[t]; }                     // busy half: descriptor load in flight
                }
            }
            const bool active = (L.tid[0] >= 0) || (L.tid[1] >= 0);
            if (!__any_sync(0xFFFFFFFFu, active)) break;
            L.run_block(P, prof, keep_ge, keep_lt, Q, q);
        }
    }
    // per-lane work counters -> global (one atomic per lane per launch)
    if (L.n_tasks) {
        atomicAdd(&P.counters[0], L.n_tasks);
        atomicAdd(&P.counters[1], L.n_cells);
    }
    if (L.n_steps) atomicAdd(&P.counters[2], L.n_steps);
    if (L.n_comp) atomicAdd(&P.counters[3], L.n_comp);
}

template <int NR>
__global__ void __launch_bounds__(SCORE_WARPS * 32, ScoreOcc<NR>::value)
k_score(const ScoreParams P) {
    score_kernel_body<NR, ScoreLane<NR>>(P);
}

template <int NR>
inline size_t score_smem_bytes() {
    return sizeof(U4) * 34 + sizeof(U2) * ((size_t)SCORE_WARPS * NIDX * ScoreLane<NR>::PROF_STRIDE);
}
#endif  // __CUDACC__

}  // namespace lsw
