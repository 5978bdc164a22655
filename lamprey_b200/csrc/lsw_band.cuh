// lsw_band.cuh -- K2, first tier: bounded-window traceback whose window never leaves the SM.
//
// Same job and same exactness argument as k_trace_fast (lsw_trace.cuh): recompute the values of a short,
// certified window of columns ending at K1's (max_row, max_col) and walk back from that cell, evaluating
// swalign's op rule lazily at the cells the walk visits (semantics: SURVEY.md Appendix A.1/A.2/A.4; replaces the
// traceback of the align() call at /root/reference/lamprey.py:563 and the hit test / split of
// /root/reference/lamprey.py:612-640).  What differs is where the window lives:
//
//   k_trace_fast writes H mod 16 of EVERY window cell to a global scratch and the walk reads it back through
//   L2/HBM with dependent loads (round 1 of this repo: 27 % of a step, DRAM 26-29 %, 3 long-scoreboard stalls per
//   issue).  But a traceback only ever looks at cells next to its own path, and the path hugs the diagonal through
//   the end cell: after k columns it stands on row  max_row - k + (#D - #I so far).  So per column only a BAND of
//   8*BW rows around that diagonal is kept -- BW 32-bit words of nibbles per (column, task) -- in SHARED memory:
//
//       band(k)  = rows  max_row - k - BL .. max_row - k - BL + 8 BW - 1        k = max_col - column
//
//   rows <= 0 are the matrix boundary (value 0), so the band is empty beyond k = max_row + 8 BW - 2 - BL and a
//   task needs at most KCOLS = NR + 8 BW - 1 - BL columns of BW words: 224 bytes per lane (two tasks) for NR = 24,
//   BW = 1 against ~1.5 KB of global scratch before.  The fill still runs over the whole certified window
//   (trace_window_fast: the values inside the band must be exact), but only the band columns are packed and stored.
//
// A walk that steps out of the band (more than 2-3 net gaps to one side for BW = 1, 6-7 for BW = 2), or that the
// lazy evaluation cannot prove, is not guessed: the candidate goes to the next tier (the same kernel with a wider
// band, then k_trace_full).
#pragma once
#include "lsw_common.cuh"
#include "lsw_score.cuh"
#include "lsw_trace.cuh"

namespace lsw {

template <int BW> struct BandGeom {
    static constexpr int NIB = 8 * BW;                       // band rows per column
    static constexpr int BL = BW == 1 ? 3 : NIB / 2 - 1;     // nibble position of the diagonal through the end cell
};

#if defined(LSW_EMU_STATS)
static unsigned long long g_band_walks[4] = {0, 0, 0, 0}, g_band_fail[4] = {0, 0, 0, 0}, g_band_iters[4] = {0, 0, 0, 0}, g_band_why[4][8] = {{0}}, g_band_out[4][2][2] = {{{0}}}, g_band_cert[4] = {0}, g_band_pruned[4] = {0}, g_dmin[4][64] = {{0}}, g_dmax[4][64] = {{0}}, g_drow[4][64] = {{0}};
static int g_cur_dmin = 0, g_cur_dmax = 0;
#endif

template <int NR, int BW>
struct BandLane {
    static_assert(NR % 4 == 0 && NR <= 64, "rows in registers");
    static_assert(BW == 1 || BW == 2, "band of 8 or 16 rows");
    static constexpr int PROF_STRIDE = NR / 2 + 1;
    static constexpr int NW = (NR + 7) / 8;                  // nibble words of a whole column of one task
    static constexpr int NIB = BandGeom<BW>::NIB;
    static constexpr int BL = BandGeom<BW>::BL;
    static constexpr int KCOLS = NR + NIB - 1 - BL;          // band columns a task can need (k = 0 .. KCOLS - 1)
    static constexpr int MAXD = 15;                          // nesting limit of run-flag evaluations
    static constexpr int WORDS_PER_LANE = KCOLS * 2 * BW;    // shared-memory words of one lane (two tasks)

    // stride between a lane's consecutive band words: the warp's words are interleaved on the device, the host emulation has one lane
#if defined(__CUDA_ARCH__)
    static constexpr int LS = 32;
#else
    static constexpr int LS = 1;
#endif

    uint32_t M[NR];

    struct Half { int ncols; int64_t code0; int row; };

    LSW_HD static uint32_t fsr(uint32_t lo, uint32_t hi, int sh) {
#if defined(__CUDA_ARCH__)
        return __funnelshift_r(lo, hi, sh);
#else
        return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
    }
    // (idx == r) ? a : b  as one compare + select.  Written in PTX on the device: the compiler otherwise turns a chain of these
    // over a register array into an indexed load from a local-memory copy of the array, and then keeps that copy up to date
    // with one move per row and column in the fill loop.
    LSW_HD static uint32_t sel_eq(int idx, int r, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        uint32_t o;
        asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\tselp.b32 %0, %3, %4, p;\n\t}" : "=r"(o) : "r"(idx), "r"(r), "r"(a), "r"(b));
        return o;
#else
        return idx == r ? a : b;
#endif
    }
    // word j of a column's nibble words, 0 outside (rows <= 0 and rows past the registers)
    LSW_HD static uint32_t pick(const uint32_t* N, int j) {
        uint32_t v = 0;
#pragma unroll
        for (int jj = 0; jj < NW; jj++) v = sel_eq(j, jj, N[jj], v);
        return v;
    }

    // One traceback over the band in shared memory.  State machine and decision table are k_trace_fast's
    // (TraceLane::Walker, decision_entry); only the operand fetch differs.
    template <bool OPS>
    struct Walker {
        const uint32_t* band;                  // this task's band words: word (k, w) at (k * 2 * BW + w) * LS
        const uint8_t* qc; const uint32_t* ref4; int rlead;
        int score, c0, row, ncols;
        int r, cl, hv, depth, retc, last, state, budget;
        const uint8_t* dtab;
        uint32_t mv, st;
        bool failed;
        uint32_t rw, rwn; int tagR;           // read codes: the aligned word of the current column and the one before it (fetched ahead)

        LSW_HD bool init(const TraceParams& P, const TaskView& v, int c0_, int ncols_, int h_, const uint32_t* band_,
                         const uint8_t* qcodes, const uint8_t* table) {
            dtab = table; failed = false;
            depth = last = state = 0; retc = 0; mv = 0; st = 0;
            band = band_ + h_ * BW * LS;
            qc = qcodes;
            const uint8_t* refw = v.ref + c0_ - 1;
            rlead = (int)(reinterpret_cast<uintptr_t>(refw) & 3u);
            ref4 = reinterpret_cast<const uint32_t*>(refw - rlead);
            score = v.score; c0 = c0_; row = v.row; ncols = ncols_;
            r = v.row; cl = ncols_; hv = v.score;
            budget = 6 * (v.row + ncols_) + 64;
            tagR = (cl + rlead) >> 2;           // (the code array is padded by 256 bytes in front: word tagR - 1 always exists)
            rw = ref4[tagR]; rwn = ref4[tagR - 1];
            if (c0_ > 0) {      // one certificate for the whole walk (trace_slack): margin of the start cell
                const int num = P.match * r - score - 1;
                const int margin = cl - r - (num >= 0 ? num / P.g : -1);
                if (margin < trace_slack(P.match, P.g) - 1) {
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
                    g_band_cert[BW]++;
#endif
                    failed = true; return false; }
            }
            return true;
        }
        LSW_HD static int nibw(uint32_t w, int p) { return (int)((w >> (4 * p)) & 15u); }
        LSW_HD int nib(const uint32_t* w, int p) const {          // nibble p of a band column (BW words)
            if (BW == 1) return nibw(w[0], p);
            return nibw(p >= 8 ? w[BW - 1] : w[0], p & 7);
        }
        LSW_HD bool step(const TraceParams& P, PathOutT<OPS>& po) {
            if (r <= 0 || cl <= 0) return false;                  // only at depth 0: the matrix boundary is zero
            const int k = ncols - cl;
            const int p0 = r - row + k + BL;                      // nibble of (r, cl) in band column k
            const bool inband = (p0 >= 1) & (p0 <= NIB - 2);
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            { const int dl = p0 - BL; if (dl < g_cur_dmin) g_cur_dmin = dl; if (dl > g_cur_dmax) g_cur_dmax = dl; if (!inband) g_drow[BW][r < 63 ? r : 63]++; }
#endif
            const int p = inband ? p0 : 1, kc = inband ? k : 0;   // (a cell outside the band fails below; keep the loads in range)
            uint32_t A[BW], B[BW];
#pragma unroll
            for (int w = 0; w < BW; w++) {
                A[w] = band[(kc * 2 * BW + w) * LS];
                B[w] = cl > 1 ? band[((kc + 1) * 2 * BW + w) * LS] : 0u;    // column 0 of a slice is the zero boundary
            }
            const int i = r - 1;
            const int off = cl + rlead;
            // the walk moves one column at a time: word tagR - 1 is already here when the column crosses into it
            if ((off >> 2) < tagR) { tagR--; rw = rwn; rwn = ref4[tagR - 1]; }
            else if ((off >> 2) > tagR) { tagR++; rwn = rw; rw = ref4[tagR]; }
            const int rb = (int)((rw >> (8 * (off & 3))) & 0xFFu);
            const int qb = qc[i];
            const int g = P.g;
            const bool inner = (i > 0) & (cl > 1);
            const int sub = (qb == rb) ? P.match : P.mismatch;
            const int hC = nib(A, p), hU = nib(A, p - 1), hL = nib(B, p + 1), hD = nib(B, p);
            const bool tieL = (cl > 1) & (((hL - g - hv) & 15) == 0);
            const bool tieU = (i > 0) & (((hU - g - hv) & 15) == 0);
            const bool tieM = inner ? (((hD + sub - hv) & 15) == 0) : (hv == sub);
            const bool s0 = state == 0;
            if (depth == 0 && s0 && hv == 0) return false;           // path ends at the first zero cell
            budget -= s0 ? 1 : 0;
            const int e = dtab[state | retc | (tieL ? 16 : 0) | (tieU ? 32 : 0) | (tieM ? 64 : 0)];
            const int res = e & 3;
            const bool pushL = (e & 4) != 0, pushU = (e & 8) != 0, push = (e & 12) != 0;
            const bool bad = !inband | (((hC - hv) & 15) != 0) | ((cl <= 1) & (c0 > 0)) | (budget < 0) |
                             ((e & 16) != 0) | (push & (depth + 1 >= MAXD));
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            if (bad) {
                const int why = !inband ? 0 : (((hC - hv) & 15) != 0) ? 1 : ((cl <= 1) & (c0 > 0)) ? 2 : (budget < 0) ? 3 : ((e & 16) != 0) ? 4 : 5;
                g_band_why[BW][why]++;
                if (!inband) g_band_out[BW][p0 < 1 ? 0 : 1][depth > 0 ? 1 : 0]++;
            }
#endif
            if (bad) { failed = true; return false; }
            const bool isret = !push & (depth > 0);
            const bool isres = !push & (depth == 0);
            const int up1 = (int)(mv & 1u);
            const int dr = push ? -(int)pushU : (isret ? up1 : -(int)(res != OPC_D));
            const int dc = push ? -(int)pushL : (isret ? 1 - up1 : -(int)(res != OPC_I));
            state = isret ? (int)(st & 3u) : 0;
            st = push ? ((st << 2) | (pushU ? 2u : 1u)) : (isret ? st >> 2 : st);
            mv = push ? ((mv << 1) | (pushU ? 1u : 0u)) : (isret ? mv >> 1 : mv);
            depth += push ? 1 : (isret ? -1 : 0);
            retc = push ? retc : ((e >> 3) & 12);
            r += dr; cl += dc;
            hv += push ? g : (isret ? -g : (res == OPC_M ? -sub : g));
            const bool ismatch = (res == OPC_M) & (sub == P.match);
            po.nrun += (isres & (res != last)) ? 1 : 0;
            po.m += (isres & ismatch) ? 1 : 0;
            po.x += (isres & !ismatch) ? 1 : 0;
            last = isres ? res : last;
            if (OPS) { if (isres && !po.push(res)) { failed = true; return false; } }
            return true;
        }
        LSW_HD bool finish(PathOutT<OPS>& po) const {
            if (failed) return false;
            if (cl == 0 && c0 > 0 && r > 0) return false;
            po.q_pos = r;
            po.r_pos = c0 + cl;
            return true;
        }
    };

    LSW_HD static void window(const TraceParams& P, const TaskView& v, int& c0, int& ncols) {
        c0 = 0; ncols = 0;
        if (v.score > 0) {
            const int W = trace_window_fast(v.row, v.score, P.match, P.g);
            c0 = v.col > W ? v.col - W : 0;
            ncols = v.col - c0;
        }
    }

    // A candidate whose score is  match * max_row  needs no DP at all: the only way to reach that value in row max_row is max_row
    // matches on the diagonal from row 0 (every gap or mismatch loses at least one point against the row's ceiling), no neighbour
    // can tie it (L - g and U - g stay below the ceiling), and the traceback stops at row 0, where the matrix is zero.  So the
    // alignment is  q_pos 0, q_end max_row, r_pos max_col - max_row, CIGAR (max_row)M, max_row matches, no mismatches -- 13 % of
    // the candidates of a 24-base primer at 8 % read error, 26 % for 16 bases; they sort next to each other (shortest windows),
    // so whole pairs skip the fill.
    LSW_HD static bool is_perfect(const TraceParams& P, const TaskView& v) { return v.score > 0 && v.score == P.match * v.row; }

    template <bool OPS>
    LSW_HD static void emit_perfect(const TraceParams& P, const TaskView& v) {
        PathOutT<OPS> po;
        po.clear();
        po.q_pos = 0; po.r_pos = v.col - v.row; po.m = v.row; po.x = 0; po.nrun = 1;
        if (OPS) for (int i = 0; i < v.row; i++) po.push(OPC_M);
        emit(P, v.t, v, po);
    }

    static constexpr uint32_t XBIAS = 254;                   // bias of the packed values: low byte = XBIAS - (errors on the best path)

    // The fill is LEXICOGRAPHIC: every packed value is  256 * H - x + XBIAS,  x = the fewest errors (mismatches + gap
    // bases, swalign's `mismatches`) over all H-optimal paths into the cell, restarted at zero cells exactly as the
    // traceback stops there.  It is the same recurrence with the scores  match 256 M, mismatch 256 X - 1, gap 256 g + 1
    // (same instruction count), byte 1 / 3 of a register is still H (XBIAS - x stays within the low byte), and the value of
    // the end cell yields x_min: the path swalign's tie rule picks is one of those H-optimal paths, so its error count
    // is >= x_min, and a candidate whose (score, x_min) can no longer pass the hit test is dropped without a walk.
    LSW_HD static int32_t lex_match(int match, int g) { return SC * (match + g) + 1; }
    LSW_HD static int32_t lex_mis(int mismatch, int g) { return SC * (mismatch + g); }

    // Fill the certified windows of the pair (ci0, ci1) of sequence q (ci1 may be -1) and store their bands.
    // band: this lane's shared-memory words (stride LS between consecutive words).
    // xmin[h]: fewest errors of an optimal path into the end cell; -1 = nothing to walk (score 0 or no candidate),
    // -2 = the recomputed end cell does not reproduce K1's score (never expected: next tier), -3 = a perfect match (emit_perfect).
    LSW_HD void fill_pair(const TraceParams& P, const U2* prof, int q, int ci0, int ci1, uint32_t* band, int* xmin) {
        Half hf[2];
        const int ci[2] = {ci0, ci1};
        int score[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            hf[h].ncols = 0; hf[h].code0 = 0; hf[h].row = 0; score[h] = 0; xmin[h] = -1;
            if (ci[h] >= 0) {
                const TaskView v = load_cand(P, ci[h], q);
                if (is_perfect(P, v)) { xmin[h] = -3; continue; }      // nothing to fill or walk: see emit_perfect
                int c0;
                window(P, v, c0, hf[h].ncols);
                hf[h].code0 = (v.ref - P.codes) + c0;
                hf[h].row = v.row; score[h] = v.score;
            }
        }
        const int n = hf[0].ncols > hf[1].ncols ? hf[0].ncols : hf[1].ncols;
        const int first[2] = {n - hf[0].ncols, n - hf[1].ncols};
        if (n <= 0) return;

        const uint32_t G2 = (uint32_t)(P.g * SC + 1) * 0x00010001u;
        const uint32_t B2 = XBIAS * 0x00010001u;
        const uint32_t FLOOR2 = G2 + B2;
#pragma unroll
        for (int r = 0; r < NR; r++) M[r] = B2;
        // read codes: aligned words.  Three words are in flight per task: the pair the current four steps are cut from and the
        // word after them, which was requested one iteration (four columns) before it is looked at -- the value loaded in an
        // iteration must not be touched in the same iteration, not even by a register move, or the warp waits for it there
        // (ncu: a quarter of the fill's stall samples sat on exactly that move).
        int64_t addr[2]; int lead[2]; uint32_t w0[2], w1[2], w2[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            // (a half without a task reads the front pad of the code array: the loads are unconditional on purpose, a predicated
            // load lands in a temporary that is then moved into the loop-carried register in the same iteration)
            const int64_t a0 = hf[h].ncols > 0 ? hf[h].code0 - first[h] : 0;
            lead[h] = (int)(a0 & 3);
            addr[h] = a0 - lead[h];
            w0[h] = *reinterpret_cast<const uint32_t*>(P.codes + addr[h]);
            w1[h] = *reinterpret_cast<const uint32_t*>(P.codes + addr[h] + 4);
            w2[h] = *reinterpret_cast<const uint32_t*>(P.codes + addr[h] + 8);
        }
#pragma unroll 1
        for (int s0 = 0; s0 < n; s0 += 4) {
            uint32_t c[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t v = prmt(w0[h], w1[h], 0x3210u + 0x1111u * (uint32_t)lead[h]);
                w0[h] = w1[h];
                w1[h] = w2[h];                                                               // loaded one iteration ago
                w2[h] = *reinterpret_cast<const uint32_t*>(P.codes + addr[h] + s0 + 12);     // looked at one iteration from now
                if (h) v |= (v & 0x04040404u) << 2;
                int nl = first[h] - s0; nl = nl < 0 ? 0 : (nl > 4 ? 4 : nl);
                const uint32_t kp = (hf[h].ncols > 0 && nl < 4) ? (0xFFFFFFFFu << (8 * nl)) : 0u;
                const uint32_t OTH = 0x01010101u * (h ? CODE_OTHER_HI : CODE_OTHER);
                c[h] = (v & kp) | (OTH & ~kp);
            }
            const uint32_t idx4 = c[0] * 4u + c[1];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (s0 + j < n) {
                    const uint32_t idx = (idx4 >> (8 * j)) & 0xFFu;
                    const U2* prow = prof + idx * PROF_STRIDE;
                    uint32_t d = B2, u = B2;
#pragma unroll
                    for (int r2 = 0; r2 < NR / 2; r2++) {
                        const U2 sv = prow[r2];
                        uint32_t mold = M[2 * r2];
                        uint32_t T = addmax2(d, sv.x, mold);
                        uint32_t hh = max3_2(T, u, FLOOR2);
                        u = hh - G2; M[2 * r2] = u; d = mold;
                        mold = M[2 * r2 + 1];
                        T = addmax2(d, sv.y, mold);
                        hh = max3_2(T, u, FLOOR2);
                        u = hh - G2; M[2 * r2 + 1] = u; d = mold;
                    }
                    const int k = n - 1 - (s0 + j);               // distance from the end column: the same for both tasks (right-aligned)
                    if (k < KCOLS) {
                        // nibbles of the whole column: H of row r is byte 1 (low task) / byte 3 (high task) of M[r]
                        uint32_t Nlo[NW], Nhi[NW];
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            uint32_t pr[4];
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                const int ra = 8 * w + 2 * kk, rb = ra + 1;
                                const uint32_t va = ra < NR ? M[ra < NR ? ra : 0] : 0u;
                                const uint32_t vb = rb < NR ? M[rb < NR ? rb : 0] : 0u;
                                pr[kk] = (va & 0x0F000F00u) | ((vb << 4) & 0xF000F000u);
                            }
                            const uint32_t p01 = prmt(pr[0], pr[1], 0x7531u), p23 = prmt(pr[2], pr[3], 0x7531u);
                            Nlo[w] = prmt(p01, p23, 0x6420u);
                            Nhi[w] = prmt(p01, p23, 0x7531u);
                        }
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int b0 = hf[h].row - k - BL - 1;        // 0-based row of the band's nibble 0 (may be negative)
                            const int wi = (b0 - (b0 & 7)) / 8;           // floor(b0 / 8)
                            const int sh = (b0 & 7) * 4;
                            const uint32_t* N = h ? Nhi : Nlo;
                            uint32_t wv[BW + 1];
#pragma unroll
                            for (int w = 0; w <= BW; w++) wv[w] = pick(N, wi + w);
#pragma unroll
                            for (int w = 0; w < BW; w++) band[((k * 2 + h) * BW + w) * LS] = fsr(wv[w], wv[w + 1], sh);
                        }
                    }
                }
            }
        }
        // the end cells: row max_row of the last column
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (hf[h].ncols <= 0) continue;
            uint32_t v = 0;
#pragma unroll
            for (int r = 0; r < NR; r++) v = sel_eq(hf[h].row - 1, r, M[r], v);
            const uint32_t half = h ? (v >> 16) : (v & 0xFFFFu);
            xmin[h] = ((int)(half >> 8) == score[h]) ? (int)XBIAS - (int)(half & 0xFFu) : -2;
        }
    }

    // The traceback of candidate record cidx (task half h of the lane whose band this is) -> hit test / result record.
    // false: the candidate must go to the next tier.
    template <bool OPS>
    LSW_HD bool walk_one(const TraceParams& P, const uint8_t* qcodes, const uint8_t* dtab, int q, int cidx, int h, const uint32_t* band) {
        const TaskView v = load_cand(P, cidx, q);
        PathOutT<OPS> po;
        po.clear();
        po.q_pos = v.row; po.r_pos = v.col;
        bool ok = true;
        if (v.score > 0) {
            int c0, ncols;
            window(P, v, c0, ncols);
            Walker<OPS> w;
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            g_cur_dmin = g_cur_dmax = 0;
#endif
            if (w.init(P, v, c0, ncols, h, band, qcodes, dtab))
                while (w.step(P, po)) {
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
                    g_band_iters[BW]++;
#endif
                }
            ok = w.finish(po);
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            g_band_walks[BW]++; if (!ok) g_band_fail[BW]++;
            if (ok) { g_dmin[BW][-g_cur_dmin < 63 ? -g_cur_dmin : 63]++; g_dmax[BW][g_cur_dmax < 63 ? g_cur_dmax : 63]++; }
#endif
        }
        if (ok) emit(P, v.t, v, po);
        return ok;
    }

    // can a candidate of sequence q with this score and at least xmin errors still pass the hit test?  (find_all only:
    // cand_ok[q][score] holds 1 + the largest error count of any passing alignment with that score, lsw_host.hpp)
    LSW_HD static bool may_pass(const TraceParams& P, int q, int score, int xmin) {
        return P.mode != 0 || xmin < (int)P.cand_ok[(size_t)q * 128 + score];
    }

    // fill + walk of one pair, one lane at a time (host emulation and the simple device path).  fb(ci, t): next tier.
    template <bool OPS, class Fallback>
    LSW_HD void run_pair(const TraceParams& P, const U2* prof, const uint8_t* qcodes, const uint8_t* dtab, int q, int ci0, int ci1,
                         uint32_t* band, Fallback fb) {
        int xmin[2];
        fill_pair(P, prof, q, ci0, ci1, band, xmin);
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int cidx = h ? ci1 : ci0;
            if (cidx < 0) continue;
            const TaskView v = load_cand(P, cidx, q);
            if (xmin[h] == -3) { emit_perfect<OPS>(P, v); continue; }
            if (xmin[h] == -2) { fb(cidx, v.t); continue; }
            if (xmin[h] >= 0 && !may_pass(P, q, v.score, xmin[h])) {
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
                g_band_pruned[BW]++;
#endif
                continue;                                   // certainly not a hit: no record, no children
            }
            if (!walk_one<OPS>(P, qcodes, dtab, q, cidx, h, band)) fb(cidx, v.t);
        }
    }
};

#if defined(__CUDACC__)
// ---- the kernel --------------------------------------------------------------------------
// Candidates arrive sorted by (sequence, window length).  A BLOCK serves one sequence at a time (its lexicographic profile
// and query codes sit in shared memory once per block); its warps claim 64 consecutive candidates of that sequence at a
// time: every lane fills the windows of one PAIR, then the candidates that survive the error-count test are dealt out again
// over the lanes of the warp for the walk -- about half of them are dropped by that test, and a lane walking its own two
// candidates one after the other left two thirds of the warp idle (ncu: 9.6 of 32 threads active in the walk loop).
// Each warp owns WORDS_PER_LANE x 32 words of band storage, word-interleaved across its lanes (conflict-free for the
// fill; the dealt-out walks read other lanes' columns, at worst two lanes share a bank).
template <int NR, int BW> struct BandSmem {
    using Lane = BandLane<NR, BW>;
    static constexpr size_t prof_bytes = sizeof(U2) * ((size_t)NIDX * Lane::PROF_STRIDE);
    static constexpr size_t seg_bytes = 2 * 256 * sizeof(int32_t);
    static constexpr size_t misc_bytes = QSTRIDE + 128 + 16;
    static constexpr size_t band_bytes = (size_t)(TRACE_THREADS / 32) * Lane::WORDS_PER_LANE * 32 * 4;
    static constexpr size_t total = prof_bytes + seg_bytes + misc_bytes + band_bytes;
};

template <int NR, int BW, bool OPS>
__global__ void __launch_bounds__(TRACE_THREADS, (NR <= 32 ? 3 : 1))
k_trace_band(const TraceParams P) {
    using Lane = BandLane<NR, BW>;
    using SM = BandSmem<NR, BW>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    U2* prof = reinterpret_cast<U2*>(smem_raw);
    int32_t* seg_lo = reinterpret_cast<int32_t*>(smem_raw + SM::prof_bytes);
    int32_t* seg_hi = seg_lo + 256;
    uint8_t* qsm = reinterpret_cast<uint8_t*>(seg_hi + 256);                           // the block's query codes
    uint8_t* dtab = qsm + QSTRIDE;
    int32_t* s_sel = reinterpret_cast<int32_t*>(dtab + 128);
    uint32_t* band_warp = reinterpret_cast<uint32_t*>(smem_raw + SM::prof_bytes + SM::seg_bytes + SM::misc_bytes) +
                          (size_t)warp * Lane::WORDS_PER_LANE * 32;
    const long long ncand = (long long)*P.ncand;
    dtab[threadIdx.x] = decision_entry((int)threadIdx.x);                              // 128 threads, 128 entries
    for (int i = threadIdx.x; i < P.n_class_seqs; i += blockDim.x) {
        const int q = P.class_seqs[i];
        seg_lo[i] = (int32_t)lower_bound_u16(P.cand_key, ncand, (uint32_t)q << 8);
        seg_hi[i] = (int32_t)lower_bound_u16(P.cand_key, ncand, ((uint32_t)q + 1) << 8);
    }
    __syncthreads();
    Lane L;
    int prefer = (int)(blockIdx.x % (unsigned)P.n_class_seqs);
    for (;;) {
        if (threadIdx.x == 0) {                                  // next sequence of the class that still has unclaimed candidates
            int found = -1;
            for (int t = 0; t < P.n_class_seqs && found < 0; t++) {
                int i = prefer + t; if (i >= P.n_class_seqs) i -= P.n_class_seqs;
                const long long hd = (long long)*((volatile unsigned long long*)&P.head[P.class_seqs[i]]);
                if (hd < (long long)(seg_hi[i] - seg_lo[i])) found = i;
            }
            *s_sel = found;
        }
        __syncthreads();
        const int sel = *s_sel;
        if (sel < 0) break;
        const int q = P.class_seqs[sel];
        const long long lo = seg_lo[sel];
        const long long cnt = seg_hi[sel] - lo;
        {
            ScoreParams SP;
            SP.qcodes = P.qcodes; SP.s_match = Lane::lex_match(P.match, P.g); SP.s_mis = Lane::lex_mis(P.mismatch, P.g);
            ScoreLane<NR>::build_profile(SP, q, P.qlen[q], reinterpret_cast<uint32_t*>(prof), threadIdx.x, TRACE_THREADS);
            if (threadIdx.x < QSTRIDE / 4)
                reinterpret_cast<uint32_t*>(qsm)[threadIdx.x] = reinterpret_cast<const uint32_t*>(P.qcodes + (size_t)q * QSTRIDE)[threadIdx.x];
        }
        __syncthreads();
        for (;;) {
            unsigned long long first = 0;
            if (lane == 0) first = atomicAdd(&P.head[q], 64ull);          // one claim per warp (64 candidates)
            const long long i = (long long)__shfl_sync(0xFFFFFFFFu, first, 0) + 2 * lane;
            const bool active = i < cnt;
            if (!__any_sync(0xFFFFFFFFu, active)) break;
            const int c0 = active ? (int)P.cand[lo + i] : -1;
            const int c1 = (active && i + 1 < cnt) ? (int)P.cand[lo + i + 1] : -1;
            if (P.all_fallback) {                                 // scoring outside the nibble walk's range: straight to the full window
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int ci = h ? c1 : c0;
                    if (ci >= 0) { const unsigned long long j = atomicAdd(P.nfallback, 1ull); P.fallback[j] = (uint32_t)P.cand_rec[ci].t; }
                }
                continue;
            }
            int xmin[2] = {-1, -1};
            if (active) L.fill_pair(P, prof, q, c0, c1, band_warp + lane, xmin);
            // what becomes of the two candidates: dropped (error count), next tier (-2), or walked
            bool walk[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ci = h ? c1 : c0;
                walk[h] = false;
                if (ci < 0) continue;
                if (xmin[h] == -3) { Lane::template emit_perfect<OPS>(P, load_cand(P, ci, q)); continue; }
                if (xmin[h] == -2) { const unsigned long long j = atomicAdd(&P.fb2_cnt[q], 1ull); P.fb2[lo + j] = (uint32_t)ci; continue; }
                walk[h] = xmin[h] < 0 || Lane::may_pass(P, q, P.cand_rec[ci].res_x & 0xFFFF, xmin[h]);
            }
            __syncwarp();                                         // the bands are read by other lanes from here on
            const unsigned m0 = __ballot_sync(0xFFFFFFFFu, walk[0]), m1 = __ballot_sync(0xFFFFFFFFu, walk[1]);
            const int n0 = __popc(m0), total = n0 + __popc(m1);
            for (int base = 0; base < total; base += 32) {
                const int j = base + lane;
                const bool mine = j < total;
                const int h = (mine && j >= n0) ? 1 : 0;
                const int src = mine ? (int)__fns(h ? m1 : m0, 0, (h ? j - n0 : j) + 1) : 0;      // the (j+1)-th walker of the warp
                const int s0 = __shfl_sync(0xFFFFFFFFu, c0, src), s1 = __shfl_sync(0xFFFFFFFFu, c1, src);
                if (mine) {
                    const int ci = h ? s1 : s0;
                    if (!L.template walk_one<OPS>(P, qsm, dtab, q, ci, h, band_warp + src)) {
                        const unsigned long long k = atomicAdd(&P.fb2_cnt[q], 1ull);
                        P.fb2[lo + k] = (uint32_t)ci;
                    }
                }
            }
            __syncwarp();                                         // ... until here: the next fill overwrites them
        }
        __syncthreads();                                          // every warp is done with this profile
        prefer = sel + 1 < P.n_class_seqs ? sel + 1 : 0;
    }
}
#endif

}  // namespace lsw
