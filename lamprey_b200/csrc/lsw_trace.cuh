// lsw_trace.cuh -- K2: bounded-window traceback of swalign.LocalAlignment.align()
//
// Replaces the traceback loop, _reduce_cigar and Alignment.__init__ of the third-party
// swalign module behind /root/reference/lamprey.py:563-566 (semantics: SURVEY.md Appendix
// A.1, A.2, A.4), and -- in find_all mode -- the hit test and the left/right split of
// findPrimerAlignments (/root/reference/lamprey.py:612-640).
//
// K1 (lsw_score.cuh) has already produced the exact (score, max_row, max_col).  K2 recomputes
// the matrix on a window of columns ending at max_col and walks back from (max_row, max_col).
// swalign's op rule for a positive cell (prefer_gap_runs=True, gap == gap_ext = -g):
//   'd' if left is a positive 'd' cell and H == L - g, else 'i' if up is a positive 'i' cell
//   and H == U - g, else 'm' if H == D + s, else 'd' if H == L - g, else 'i'.
//
// Two kernels:
//
//  k_trace_fast  (every candidate)  VALUES ONLY on a short window W1 = 3*max_row - score + 8
//      columns, one byte per cell in scratch; the walk then evaluates the op rule LAZILY, only
//      at the cells it visits, and proves each decision exact with a per-cell certificate:
//        * a windowed value Hw(r,c) can only be too small if the true optimal path into the cell
//          starts left of the window; such a path spans at most r + (M*r - h)/g columns, so
//          c - c0 >= r + (M*r - Hw - 1)/g  =>  Hw == H                      ("certified");
//        * a windowed TIE (H == L - g etc.) is always a true tie (Hw <= H <= the tying value),
//          and makes the neighbour's windowed value exact; only a windowed NON-tie needs the
//          neighbour certified;
//        * the run flags ("left is a 'd' cell") are obtained by evaluating the same rule at
//          the neighbour, recursively, within an evaluation budget.
//      Anything it cannot prove (uncertified neighbour, budget exhausted, window too short) is
//      pushed to a fallback list -- never guessed.
//
//  k_trace_full  (fallback list)  the op rule for every cell of a window that is long enough
//      to be exact a priori.  With S = Q + M*Q/g (upper bound of a path's column span) and
//      C = M*Q/g (upper bound of a run-flag dependency chain):  W = 2S + C + 4  (the 8Q of
//      SURVEY.md Appendix A.4 for +2/-1/-1).  A window starting at the slice start is the true
//      boundary and needs no argument.
#pragma once
#include "lsw_common.cuh"
#include "../../include/lsw.h"

namespace lsw {

LSW_HD int trace_window_full(int Q, int match, int g) {
    const int dmax = (match * Q) / g;          // >= number of 'D' moves on any positive-prefix path
    const int span = Q + dmax;
    return 2 * span + dmax + 4;
}
LSW_HD int trace_window_fast(int max_row, int score, int match, int g) {
    const int w = max_row + (match * max_row - score) / g + 8;
    return w;
}
LSW_HD int trace_window_fast_max(int Q, int match, int g) { return Q + (match * Q) / g + 8; }

enum { OPC_NONE = 0, OPC_M = 1, OPC_I = 2, OPC_D = 3, OPC_FAIL = -1 };
constexpr int MAX_PATH_WORDS = 16;             // 2 bits per op; a path has < Q + M*Q/g <= 3*63 ops ... 256 ops of room

struct PathOut {
    int q_pos, r_pos, m, x, nrun, nops;
    uint32_t ops[MAX_PATH_WORDS];              // ops in walk order (alignment end first)
    LSW_HD void clear() {
        q_pos = r_pos = m = x = nrun = nops = 0;
#pragma unroll
        for (int i = 0; i < MAX_PATH_WORDS; i++) ops[i] = 0;
    }
    LSW_HD bool push(int op) {
        if (nops >= MAX_PATH_WORDS * 16) return false;
        ops[nops >> 4] |= (uint32_t)op << (2 * (nops & 15));
        nops++;
        return true;
    }
    LSW_HD int get(int i) const { return (int)((ops[i >> 4] >> (2 * (i & 15))) & 3u); }
};

struct TaskView {
    int rd, a, b, q, Q, score, row, col;
    const uint8_t* ref;      // ref[c - 1] is the read code of DP column c (slice relative)
    const uint8_t* qc;       // query codes
};

LSW_HD TaskView load_task(const TraceParams& P, int t) {
    TaskView v;
    v.rd = P.t.read[t]; v.a = P.t.a[t]; v.b = P.t.b[t]; v.q = P.t.seq[t];
    const int2 rs = P.res[t];
    v.score = rs.x & 0xFFFF; v.row = rs.x >> 16; v.col = rs.y;
    v.Q = P.qlen[v.q];
    v.ref = P.codes + P.roff[v.rd] + v.a;
    v.qc = P.qcodes + (size_t)v.q * QSTRIDE;
    return v;
}

// Results of one task -> caller-visible records (both kernels end here).
LSW_HD void emit(const TraceParams& P, int t, const TaskView& v, const PathOut& po) {
    if (P.mode == 1) {
        lsw_result* out = reinterpret_cast<lsw_result*>(P.results) + P.t.orig[t];
        out->score = v.score;
        out->q_pos = po.q_pos; out->q_end = v.row;
        out->r_pos = po.r_pos; out->r_end = v.col;
        out->matches = po.m; out->mismatches = po.x;
        out->n_cigar = po.nrun;
        long long off = 0;
        if (P.cigar && po.nrun > 0) {
            off = (long long)atomic_add_ull(P.ncigar, (unsigned long long)po.nrun);
            if (off + po.nrun <= P.cigar_cap) {
                int k = 0, last = 0, cnt = 0;
                for (int i = po.nops - 1; i >= 0; i--) {          // forward (alignment start first)
                    const int op = po.get(i);
                    if (op != last) {
                        if (last) P.cigar[off + k++] = ((uint32_t)cnt << 2) | (uint32_t)last;
                        last = op; cnt = 0;
                    }
                    cnt++;
                }
                if (last) P.cigar[off + k++] = ((uint32_t)cnt << 2) | (uint32_t)last;
            }
        }
        out->cigar_off = off;
        return;
    }
    // ---- find_all mode: hit test and split (lamprey.py:612-640) ---------------------------
    const int tot = po.m + po.x;
    const int span = v.col - po.r_pos;
    // identity is float(m)/(m+x) (Python double division), or the int 0 when nothing aligned
    const double identity = tot > 0 ? (double)po.m / (double)tot : 0.0;
    if (identity >= P.thr && (double)span >= (double)v.Q * P.thr) {
        const unsigned long long hi = atomic_add_ull(P.nhits, 1ull);
        if ((long long)hi < P.hits_cap) {
            lsw_hit* h = reinterpret_cast<lsw_hit*>(P.hits) + hi;
            h->read = v.rd; h->start = v.a + po.r_pos; h->end = v.a + v.col;
            h->query = (int16_t)v.q; h->depth = (int16_t)P.round;
            h->matches = (int16_t)po.m; h->mismatches = (int16_t)po.x;
            h->score = (int16_t)v.score; h->q_pos = (int16_t)po.q_pos; h->q_end = (int16_t)v.row;
            h->reserved = 0; h->reserved2 = 0;
        }
        if (po.r_pos >= v.Q) {                       // len(left_seq) >= len(primer)
            const int64_t j = 2 * (int64_t)t;
            P.child.read[j] = v.rd; P.child.a[j] = v.a; P.child.b[j] = v.a + po.r_pos; P.child.seq[j] = (int16_t)v.q;
            P.child_flag[j] = 1;
        }
        if (v.b - (v.a + v.col) >= v.Q) {            // len(right_seq) >= len(primer)
            const int64_t j = 2 * (int64_t)t + 1;
            P.child.read[j] = v.rd; P.child.a[j] = v.a + v.col; P.child.b[j] = v.b; P.child.seq[j] = (int16_t)v.q;
            P.child_flag[j] = 1;
        }
    }
}

// ======================= full-window kernel (a-priori exact) =================================
template <int NR>
LSW_HD void trace_full_task(const TraceParams& P, const int t, const int64_t slot) {
    static_assert(NR % 16 == 0 && NR <= 64, "one 32-bit word of 2-bit ops per 16 rows");
    constexpr int WORDS = NR / 16;
    const TaskView v = load_task(P, t);
    PathOut po;
    po.clear();
    po.q_pos = v.row; po.r_pos = v.col;

    if (v.score > 0) {
        const int W = trace_window_full(v.Q, P.match, P.g);
        const int c0 = v.col > W ? v.col - W : 0;
        const int ncols = v.col - c0;

        uint32_t qw[NR / 4];                 // query codes, 4 per word; rows >= Q hold 0xFF and never match
#pragma unroll
        for (int i = 0; i < NR / 4; i++) qw[i] = reinterpret_cast<const uint32_t*>(v.qc)[i];
        int H[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) H[r] = 0;
        unsigned long long dprev = 0;        // bit r: cell (r, previous column) is a positive 'd' cell

        for (int c = 1; c <= ncols; c++) {
            const int bc = v.ref[c0 + c - 1];
            int diag = 0, up = 0;
            bool up_i = false;
            unsigned long long dcur = 0;
            uint32_t opw[WORDS];
#pragma unroll
            for (int w = 0; w < WORDS; w++) opw[w] = 0;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int qb = (int)((qw[r >> 2] >> (8 * (r & 3))) & 0xFF);
                const int s = (qb == bc) ? P.match : P.mismatch;
                const int mm = diag + s;
                const int dl = H[r] - P.g;
                const int iu = up - P.g;
                int cell = mm > dl ? mm : dl;
                cell = cell > iu ? cell : iu;
                cell = cell > 0 ? cell : 0;
                const bool left_d = (dprev >> r) & 1ull;
                int op;
                if (cell <= 0) op = OPC_NONE;
                else if (left_d && cell == dl) op = OPC_D;
                else if (up_i && cell == iu) op = OPC_I;
                else if (cell == mm) op = OPC_M;
                else if (cell == dl) op = OPC_D;
                else op = OPC_I;
                diag = H[r];
                H[r] = cell;
                up = cell;
                up_i = (op == OPC_I);
                if (op == OPC_D) dcur |= 1ull << r;
                opw[r >> 4] |= (uint32_t)op << (2 * (r & 15));
            }
            dprev = dcur;
#pragma unroll
            for (int w = 0; w < WORDS; w++)
                P.scratch[((int64_t)(c - 1) * WORDS + w) * P.scratch_stride + slot] = opw[w];
        }

        // walk back until the first non-positive cell (or the matrix boundary, whose cells are 0)
        int r = v.row, c = ncols, last = 0;
        bool ok = true;
        while (r > 0 && c > 0) {
            const uint32_t wv = P.scratch[((int64_t)(c - 1) * WORDS + ((r - 1) >> 4)) * P.scratch_stride + slot];
            const int op = (int)((wv >> (2 * ((r - 1) & 15))) & 3u);
            if (op == OPC_NONE) break;
            if (op != last) { po.nrun++; last = op; }
            ok = ok && po.push(op);
            if (op == OPC_M) {
                if (v.qc[r - 1] == v.ref[c0 + c - 1]) po.m++; else po.x++;
                r--; c--;
            } else if (op == OPC_I) { po.x++; r--; }
            else { po.x++; c--; }
        }
        if (!ok || (c == 0 && c0 > 0 && r > 0)) atomicExch_or_store(P.errflag, 1);   // window bound violated: never expected
        po.q_pos = r;
        po.r_pos = c0 + c;
    }
    emit(P, t, v, po);
}

// ======================= short-window kernel with lazy, certified op evaluation ==============
struct FastWalk {
    const uint32_t* scr; int64_t stride, slot;
    int words;               // 32-bit words per window column (4 rows per word)
    int c0;                  // DP column of the zero boundary left of the window (0 = slice start)
    const uint8_t* qc; const uint8_t* ref;
    int match, mismatch, g;
    int budget;

    LSW_HD int H(int r, int cl) const {          // windowed value; r in [0, Q], cl in [0, ncols]
        if (r <= 0 || cl <= 0) return 0;
        const uint32_t w = scr[((int64_t)(cl - 1) * words + ((r - 1) >> 2)) * stride + slot];
        return (int)((w >> (8 * ((r - 1) & 3))) & 0xFFu);
    }
    LSW_HD bool cert(int r, int cl, int h) const {
        if (c0 == 0 || r <= 0) return true;      // true boundary / row 0
        const int num = match * r - h - 1;
        const int need = r + (num >= 0 ? num / g : -1);
        return cl >= need;
    }
    // op of the positive cell (r0, cl0) whose windowed value is known to be exact; OPC_FAIL = cannot
    // prove.  The run flags need the op of the left / up neighbour, found by the same rule: an
    // explicit frame stack instead of device recursion.  state 0 = fresh, 1 = waiting for the left
    // neighbour's op, 2 = waiting for the up neighbour's op.
    static constexpr int MAXD = 24;
    LSW_HD int op_at(int r0, int cl0) {
        int fr[MAXD], fc[MAXD], fs[MAXD];
        int sp = 0, ret = OPC_NONE;
        fr[0] = r0; fc[0] = cl0; fs[0] = 0;
        while (sp >= 0) {
            const int r = fr[sp], cl = fc[sp], st = fs[sp];
            const int h = H(r, cl);
            const int hl = H(r, cl - 1), hu = H(r - 1, cl), hd = H(r - 1, cl - 1);
            const int s = (qc[r - 1] == ref[c0 + cl - 1]) ? match : mismatch;
            const bool tieL = (h == hl - g), tieU = (h == hu - g), tieM = (h == hd + s);
            if (st == 0) {
                if (--budget < 0) return OPC_FAIL;
                // a windowed non-tie is only trustworthy if that neighbour's value is certified
                if (!tieL && !cert(r, cl - 1, hl)) return OPC_FAIL;
                if (!tieU && !cert(r - 1, cl, hu)) return OPC_FAIL;
                if (!tieM && !cert(r - 1, cl - 1, hd)) return OPC_FAIL;
                if (!tieL && !tieU && !tieM) return OPC_FAIL;     // cannot happen for an exact positive cell
                if (tieL) {
                    if (!tieU && !tieM) { ret = OPC_D; sp--; continue; }
                    if (sp + 1 >= MAXD) return OPC_FAIL;
                    fs[sp] = 1;                                  // left value is h + g > 0 and exact (tie)
                    sp++; fr[sp] = r; fc[sp] = cl - 1; fs[sp] = 0;
                    continue;
                }
            }
            if (st == 1 && ret == OPC_D) { sp--; continue; }      // left is a 'd' cell: d-run (ret stays D)
            if (st != 2 && tieU) {
                if (!tieM && !tieL) { ret = OPC_I; sp--; continue; }
                if (sp + 1 >= MAXD) return OPC_FAIL;
                fs[sp] = 2;                                      // up value is h + g > 0 and exact (tie)
                sp++; fr[sp] = r - 1; fc[sp] = cl; fs[sp] = 0;
                continue;
            }
            if (st == 2 && ret == OPC_I) { sp--; continue; }      // up is an 'i' cell: i-run (ret stays I)
            ret = tieM ? OPC_M : (tieL ? OPC_D : OPC_I);
            sp--;
        }
        return ret;
    }
};

// returns true if the task was resolved; false = push it to the fallback list
template <int NR>
LSW_HD bool trace_fast_task(const TraceParams& P, const int t, const int64_t slot) {
    static_assert(NR % 4 == 0 && NR <= 64, "four 8-bit values per 32-bit word");
    constexpr int WORDS = NR / 4;
    const TaskView v = load_task(P, t);
    PathOut po;
    po.clear();
    po.q_pos = v.row; po.r_pos = v.col;
    if (v.score <= 0) { emit(P, t, v, po); return true; }

    const int W = trace_window_fast(v.row, v.score, P.match, P.g);
    const int c0 = v.col > W ? v.col - W : 0;
    const int ncols = v.col - c0;

    uint32_t qw[NR / 4];
#pragma unroll
    for (int i = 0; i < NR / 4; i++) qw[i] = reinterpret_cast<const uint32_t*>(v.qc)[i];
    int Hm[NR];                                   // H - g of the previous column
#pragma unroll
    for (int r = 0; r < NR; r++) Hm[r] = -P.g;

    for (int c = 1; c <= ncols; c++) {
        const int bc = v.ref[c0 + c - 1];
        int diag = 0, upm = -P.g;                 // row 0: H = 0
#pragma unroll
        for (int r4 = 0; r4 < NR / 4; r4++) {
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = r4 * 4 + k;
                const int qb = (int)((qw[r4] >> (8 * k)) & 0xFF);
                const int s = (qb == bc) ? P.match : P.mismatch;
                const int cell = __vimax3_s32_relu(diag + s, Hm[r], upm);
                diag = Hm[r] + P.g;
                Hm[r] = cell - P.g;
                upm = Hm[r];
                packed |= (uint32_t)cell << (8 * k);
            }
            P.scratch[((int64_t)(c - 1) * WORDS + r4) * P.scratch_stride + slot] = packed;
        }
    }

    FastWalk fw;
    fw.scr = P.scratch; fw.stride = P.scratch_stride; fw.slot = slot; fw.words = WORDS; fw.c0 = c0;
    fw.qc = v.qc; fw.ref = v.ref; fw.match = P.match; fw.mismatch = P.mismatch; fw.g = P.g;
    fw.budget = 4 * (v.row + ncols) + 64;

    int r = v.row, cl = ncols, last = 0;
    if (fw.H(r, cl) != v.score) return false;     // the window is too short to reproduce the best cell
    while (r > 0 && cl > 0) {
        // every cell on the path is exact: it was reached through a tie from an exact cell
        if (fw.H(r, cl) == 0) break;
        const int op = fw.op_at(r, cl);
        if (op == OPC_FAIL) return false;
        if (op != last) { po.nrun++; last = op; }
        if (!po.push(op)) return false;
        if (op == OPC_M) {
            if (v.qc[r - 1] == v.ref[c0 + cl - 1]) po.m++; else po.x++;
            r--; cl--;
        } else if (op == OPC_I) { po.x++; r--; }
        else { po.x++; cl--; }
    }
    if (cl == 0 && c0 > 0 && r > 0) return false; // ran into the artificial boundary
    po.q_pos = r;
    po.r_pos = c0 + cl;
    emit(P, t, v, po);
    return true;
}

#if defined(__CUDACC__)
constexpr int TRACE_THREADS = 128;

template <int NR>
__global__ void __launch_bounds__(TRACE_THREADS)
k_trace_fast(const TraceParams P) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = *P.ncand;
    for (unsigned long long i = (unsigned long long)slot; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const int t = (int)P.cand[i];
        if (!trace_fast_task<NR>(P, t, slot)) {
            const unsigned long long j = atomicAdd(P.nfallback, 1ull);
            P.fallback[j] = (uint32_t)t;
        }
    }
}

template <int NR>
__global__ void __launch_bounds__(TRACE_THREADS)
k_trace_full(const TraceParams P) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = *P.nfallback;
    for (unsigned long long i = (unsigned long long)slot; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        trace_full_task<NR>(P, (int)P.fallback[i], slot);
}
#endif

}  // namespace lsw
