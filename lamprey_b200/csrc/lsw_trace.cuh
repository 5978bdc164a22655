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
//  k_trace_fast  (every candidate)  VALUES ONLY on a short window of W1 = max_row + (M max_row - score)/g + SLACK
//      columns (trace_window_fast), H modulo 16 per cell in scratch; the walk then evaluates the op rule
//      LAZILY, only at the cells it visits, and every decision it takes is exact:
//        * a windowed value Hw(r,c) can only be too small if the true optimal path into the cell
//          starts left of the window; such a path spans at most r + (M*r - h)/g columns, so
//          c - c0 >= r + (M*r - Hw - 1)/g  =>  Hw == H                      ("certified");
//        * the window is long enough that the start cell's certificate covers every cell the walk
//          can consult (see trace_slack), so it is checked once, at the start;
//        * the walker knows the value of the cell it stands on and neighbouring values differ from
//          it by less than 16, so the stored nibbles decide every tie exactly (nibble_walk_ok);
//        * the run flags ("left is a 'd' cell") are obtained by evaluating the same rule at
//          the neighbour, recursively, within an evaluation budget.
//      Anything it cannot prove (certificate fails, budget exhausted, window too short) is
//      pushed to a fallback list -- never guessed.
//
//  k_trace_full  (fallback list)  the op rule for every cell of a window that is long enough
//      to be exact a priori.  With S = Q + M*Q/g (upper bound of a path's column span) and
//      C = M*Q/g (upper bound of a run-flag dependency chain):  W = 2S + C + 4  (the 8Q of
//      SURVEY.md Appendix A.4 for +2/-1/-1).  A window starting at the slice start is the true
//      boundary and needs no argument.
#pragma once
#include "lsw_common.cuh"
#include "lsw_score.cuh"
#include "../../include/lsw.h"
#include "lsw_chain.cuh"

namespace lsw {

constexpr int TRACE_THREADS = 128;       // block size of both kernels; also the lane stride of the fast kernel's scratch rows

LSW_HD int trace_window_full(int Q, int match, int g) {
    const int dmax = (match * Q) / g;          // >= number of 'D' moves on any positive-prefix path
    const int span = Q + dmax;
    return 2 * span + dmax + 4;
}
// Short window of the certified walk.  Define the margin of a cell  m(r, c, h) = (c - c0) - r - (M r - h - 1)/g;
// m >= 0 certifies its windowed value (see above).  Walking back along a path the margin never decreases
// (match: +0, mismatch: +M+|X|, 'i': +1+M+g.., 'd': +g-1 >= 0), a tying neighbour has a larger value and
// hence a margin no smaller, and a non-tying neighbour differs by at most M+g in value, i.e. its margin is
// at most (M+g)/g + 1 smaller.  With  W1 = max_row + (M max_row - score)/g + SLACK,  SLACK = (M+g)/g + 5,
// the start cell has margin >= SLACK - 1, so EVERY cell the walk consults is certified: the walk checks the
// start margin once instead of three certificates per step.  (For +2/-1/-1: SLACK = 8.)
LSW_HD int trace_slack(int match, int g) { return (match + g) / g + 5; }
LSW_HD int trace_window_fast(int max_row, int score, int match, int g) {
    return max_row + (match * max_row - score) / g + trace_slack(match, g);
}
// The short-window kernel stores H modulo 16.  For exact cells H(r,c) - H(r,c-1) <= M + g, H(r,c) - H(r-1,c) <= M + g and
// H(r,c) - H(r-1,c-1) <= M (drop the last column / row of the best alignment ending in (r,c)), so the three candidates
// hl - g, hu - g, hd + s lie in [H - (M + 2g), H] and [H - (M + |X|), H]: modulo 16 decides equality with H exactly when both
// spans stay below 16.  Other scorings take the full-window kernel for every candidate.
LSW_HD bool nibble_walk_ok(int match, int mismatch, int g) { return match + 2 * g <= 15 && match - mismatch <= 15; }
LSW_HD int trace_window_fast_max(int Q, int match, int g) { return Q + (match * Q) / g + trace_slack(match, g); }

enum { OPC_NONE = 0, OPC_M = 1, OPC_I = 2, OPC_D = 3, OPC_FAIL = -1 };
constexpr int MAX_PATH_WORDS = 16;             // 2 bits per op; a path has < Q + M*Q/g <= 3*63 ops ... 256 ops of room

// OPS = false (find_all): only the counts and end points are needed, the op list is never stored
template <bool OPS>
struct PathOutT {
    static constexpr int WORDS = OPS ? MAX_PATH_WORDS : 1;
    int q_pos, r_pos, m, x, nrun, nops;
    uint32_t ops[WORDS];                       // ops in walk order (alignment end first)
    LSW_HD void clear() {
        q_pos = r_pos = m = x = nrun = nops = 0;
#pragma unroll
        for (int i = 0; i < WORDS; i++) ops[i] = 0;
    }
    LSW_HD bool push(int op) {
        if (!OPS) return true;
        if (nops >= MAX_PATH_WORDS * 16) return false;
        ops[nops >> 4] |= (uint32_t)op << (2 * (nops & 15));
        nops++;
        return true;
    }
    LSW_HD int get(int i) const { return (int)((ops[i >> 4] >> (2 * (i & 15))) & 3u); }
};
using PathOut = PathOutT<true>;

struct TaskView {
    int t, rd, a, b, q, Q, score, row, col;
    const uint8_t* ref;      // ref[c - 1] is the read code of DP column c (slice relative)
    const uint8_t* qc;       // query codes
};

LSW_HD TaskView load_task(const TraceParams& P, int t) {
    TaskView v;
    v.t = t;
    v.rd = P.t.read[t]; v.a = P.t.a[t]; v.b = P.t.b[t]; v.q = P.t.seq[t];
    const int2 rs = P.res[t];
    v.score = rs.x & 0xFFFF; v.row = rs.x >> 16; v.col = rs.y;
    v.Q = P.qlen[v.q];
    v.ref = P.codes + P.roff[v.rd] + v.a;
    v.qc = P.qcodes + (size_t)v.q * QSTRIDE;
    return v;
}

// the same view from a candidate record (k_trace_fast): one 32-byte load, no dependent gathers
LSW_HD TaskView load_cand(const TraceParams& P, int ci, int q) {
    const int4* src = reinterpret_cast<const int4*>(P.cand_rec + ci);
    const int4 lo = src[0], hi = src[1];
    TaskView v;
    const int64_t p0 = (int64_t)(((unsigned long long)(uint32_t)lo.y << 32) | (unsigned long long)(uint32_t)lo.x);
    v.rd = lo.w; v.a = hi.x; v.b = hi.x + lo.z; v.q = q;
    v.score = hi.y & 0xFFFF; v.row = hi.y >> 16; v.col = hi.z; v.t = hi.w;
    v.Q = P.qlen[q];
    v.ref = P.codes + p0;
    v.qc = P.qcodes + (size_t)q * QSTRIDE;
    return v;
}

// Results of one task -> caller-visible records (both kernels end here).
template <bool OPS>
LSW_HD void emit(const TraceParams& P, int t, const TaskView& v, const PathOutT<OPS>& po) {
    if (OPS && P.mode == 1) {
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
        const unsigned long long hi = append_slot(P.nhits);
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
// Packed like K1: one LANE owns a PAIR of candidate tasks of the same sequence (low / high s16 half),
// all lanes of a warp serve one sequence whose profile sits in shared memory.  Per pair:
//   fill   n = max(window lengths) columns of VALUES with the K1 recurrence (2 DPX + 1 add per two
//          cells, no arg-max key), both windows right-aligned on their own max_col; after every
//          column the low NIBBLE of every H of both tasks goes to scratch (the traffic to and from the
//          scratch bounds this kernel) in 32-bit UNITS of eight consecutive rows, unit u = rows 7u .. 7u+7
//          (neighbouring units share a row, so a cell and the cell above it are always in one unit);
//          one 16-byte store carries two units of both tasks;
//   walk   each task: a flat state machine, one cell evaluation per iteration, that applies the op
//          rule lazily and proves every decision (see the file header).  The walker knows the VALUE of
//          the cell it stands on (the score at the start; every move changes it by a known amount), and
//          neighbouring values differ from it by less than 16 (see nibble_walk_ok), so the nibbles decide
//          every tie exactly.  A cell needs the unit of its own column and of the column to the left: the
//          walker keeps both and, because it moves one cell at a time, loads about one new unit per step.
// swalign's op rule for one cell evaluation of the walk, tabulated.  idx: bits 0-1 state (0 fresh, 1 left's op known, 2 up's op
// known), bit 2 / 3 the known op is 'd' / 'i', bits 4-6 the ties with left / up / diagonal.  Entry: bits 0-1 the cell's op
// (0: not decided yet), bit 2 / 3 descend to the left / upper neighbour first, bit 4 no tie at all (never, for an exact cell),
// bit 5 / 6 the op is 'd' / 'i' (what the asking cell will look at).
LSW_HD uint8_t decision_entry(int idx) {
    const int state = idx & 3;
    const bool retD = (idx & 4) != 0, retI = (idx & 8) != 0, tieL = (idx & 16) != 0, tieU = (idx & 32) != 0, tieM = (idx & 64) != 0;
    const bool s0 = state == 0, s1 = state == 1;
    const bool pushL = s0 && tieL && (tieU || tieM);          // need left's op (the left value is exact because of the tie)
    const bool pushU = (s0 && !tieL && tieU && tieM) || (s1 && !retD && tieU);
    const bool notie = s0 && !tieL && !tieU && !tieM;
    const int res0 = tieL ? OPC_D : (tieU ? OPC_I : OPC_M);
    const int res1 = retD ? OPC_D : (tieM ? OPC_M : OPC_D);
    const int res2 = retI ? OPC_I : (tieM ? OPC_M : (tieL ? OPC_D : OPC_I));
    const int res = (pushL || pushU) ? 0 : (s0 ? res0 : (s1 ? res1 : res2));
    return (uint8_t)(res | (pushL ? 4 : 0) | (pushU ? 8 : 0) | (notie ? 16 : 0) | (res == OPC_D ? 32 : 0) | (res == OPC_I ? 64 : 0));
}

#if defined(LSW_EMU_STATS)
static unsigned long long g_walk_iters = 0, g_walks = 0, g_fill_steps = 0;
static unsigned long long g_extra_hist[64] = {0}, g_ncols_sum = 0;   // columns a walk touched beyond max_row; window columns
static int g_min_cl = 0;
#endif

template <int NR>
struct TraceLane {
    static_assert(NR % 4 == 0 && NR <= 64, "rows in registers");
    static constexpr int PROF_STRIDE = NR / 2 + 1;
    static constexpr int UNITS = (NR + 5) / 7;    // 32-bit units per window column and task: ceil((NR - 1) / 7)
    static constexpr int UPAIRS = (UNITS + 1) / 2; // 16-byte scratch elements per window column and lane
    static constexpr int MAXD = 15;               // nesting limit of run-flag evaluations (2 state bits each)

    uint32_t M[NR];

    // what the fill needs to know about one task of the pair
    struct Half {
        int ncols, first;      // window columns, first live step of the shared loop (ncols == 0: nothing to fill)
        int64_t code0;         // offset in the code array of the read code of shared step 0
    };

    // window of a candidate: zero-boundary column c0 and number of columns
    LSW_HD static void window(const TraceParams& P, const TaskView& v, int& c0, int& ncols) {
        c0 = 0; ncols = 0;
        if (v.score > 0) {
            const int W = trace_window_fast(v.row, v.score, P.match, P.g);
            c0 = v.col > W ? v.col - W : 0;
            ncols = v.col - c0;
        }
    }

    LSW_HD static int unit_of(int i) { return i > 0 ? ((i - 1) * 37) >> 8 : 0; }     // (i - 1) / 7 for i <= 63
    LSW_HD static int nib(uint32_t v, int k) { return (int)((v >> (4 * k)) & 15u); }

    // One traceback.  Kept small on purpose: the walk is a chain of dependent loads, it is the resident warps that
    // hide them, and a walker that does not fit the register budget pays for every step in local-memory traffic.
    template <bool OPS>
    struct Walker {
        // constants of the task
        const uint32_t* scr0; uint32_t base, colbase;   // scratch array, word index of the lane's first word in it
        const uint8_t* qc; const uint32_t* ref4; int rlead;
        int score, c0;
        // state
        int r, cl, hv, depth, retc, last, state, budget;  // hv: H of the cell (r, cl); retc: 4 / 8 if the op just returned is 'd' / 'i'
        const uint8_t* dtab;
        uint32_t mv;                  // bit k: move that entered nesting level k (0 = left, 1 = up)
        uint32_t st;                  // 2 bits per level below the current one: 1 waiting for left's op, 2 for up's
        bool failed;
        // operands: the units of column cl (A) and cl - 1 (B) that hold rows r-1 and r-2, four read codes
        uint32_t A, B, C; int tagA, tagB, tagC;   // tag = 16 * column + unit, -1 = empty; C: lookahead unit
        uint32_t rw; int tagR;

        // false: nothing to walk (or the start certificate fails: see `failed`)
        LSW_HD bool init(const TraceParams& P, const TaskView& v, int c0_, int ncols, int first, int h, int64_t slot,
                         const uint8_t* qcodes, const uint8_t* table) {
            dtab = table;
            failed = false;
            depth = last = state = 0; retc = 0; mv = 0; st = 0;
            A = B = C = 0u; tagA = tagB = tagC = tagR = -1; rw = 0;
            scr0 = P.scratch;                                     // 32-bit index arithmetic: the host keeps the scratch under 2^31 words
            base = 4u * (uint32_t)slot + (uint32_t)h;
            colbase = (uint32_t)(first - 1);
            qc = qcodes;
            // ref[c0 + cl - 1] is the read code of window column cl; read through aligned words (the code array is padded)
            const uint8_t* refw = v.ref + c0_ - 1;
            rlead = (int)(reinterpret_cast<uintptr_t>(refw) & 3u);
            ref4 = reinterpret_cast<const uint32_t*>(refw - rlead);
            score = v.score; c0 = c0_;
            r = v.row; cl = ncols; hv = v.score;
            budget = 6 * (v.row + ncols) + 64;
            if (c0_ > 0) {
                // one certificate for the whole walk (see trace_slack): margin of the start cell
                const int num = P.match * r - score - 1;
                const int margin = cl - r - (num >= 0 ? num / P.g : -1);
                if (margin < trace_slack(P.match, P.g) - 1) { failed = true; return false; }
            }
            return true;
        }
        // One cell evaluation; false = the walk is over (path complete, or failed).  Written without data-dependent
        // branches (selects only, except for the rare exits): the lanes of a warp are at different points of the
        // state machine, and branches would make the warp run every arm one after the other.
        LSW_HD bool step(const TraceParams& P, PathOutT<OPS>& po) {
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            g_walk_iters++;
#endif
            if (r <= 0 || cl <= 0) return false;                  // only at depth 0: the matrix boundary is zero
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            if (cl - 1 < g_min_cl) g_min_cl = cl - 1;            // columns cl and cl - 1 are read
#endif
            const int i = r - 1, u = unit_of(i);
            {   // operands: from the cached units, else from scratch
                const int tA = 16 * cl + u, tB = cl > 1 ? tA - 16 : -2;
                const bool hitA = (tagA == tA) | (tagB == tA);      // (the lookahead unit is two columns away: never this cell's)
                const bool hitB = (tagA == tB) | (tagB == tB) | (tagC == tB) | (tB < 0);
                uint32_t nA = tagA == tA ? A : B;
                uint32_t nB = tagB == tB ? B : (tagA == tB ? A : C);
                // word (column c, unit v) of this lane: base + (c * UPAIRS + v / 2) * COLW + 2 * (v % 2); the fast kernel's
                // scratch rows are TRACE_THREADS lanes of 16 bytes
                constexpr uint32_t COLW = 4u * 128u;
                const uint32_t wA = base + ((colbase + (uint32_t)cl) * UPAIRS + ((uint32_t)u >> 1)) * COLW + 2u * ((uint32_t)u & 1u);
                if (!hitA) nA = scr0[wA];
                if (!hitB) nB = scr0[wA - UPAIRS * COLW];
                A = nA; B = nB; tagA = tA; tagB = tB;
                // lookahead: most steps are diagonal, and the diagonal successor (r-1, cl-1) takes its left-hand unit
                // from column cl-2; that load is issued now and consumed an iteration later
                const int u2 = unit_of(i > 0 ? i - 1 : 0);
                const int tC = 16 * (cl - 2) + u2;
                if (cl > 2 && tC != tagC) {
                    C = scr0[base + ((colbase + (uint32_t)cl - 2u) * UPAIRS + ((uint32_t)u2 >> 1)) * COLW + 2u * ((uint32_t)u2 & 1u)];
                    tagC = tC;
                }
            }
            const int off = cl + rlead;
            if ((off >> 2) != tagR) { tagR = off >> 2; rw = ref4[tagR]; }
            const int rb = (int)((rw >> (8 * (off & 3))) & 0xFFu);
            const int qb = qc[i];
            const int pc = i - 7 * u;
            const int g = P.g;
            // ties from the nibbles: a candidate (hl - g, hu - g, hd + sub) never exceeds hv and is less than 16 below it.
            // Cells on the matrix boundary are handled with their true value 0.
            const bool inner = (i > 0) & (cl > 1);
            const int sub = (qb == rb) ? P.match : P.mismatch;
            const bool tieL = (cl > 1) & (((nib(B, pc) - g - hv) & 15) == 0);
            const int pu1 = pc > 0 ? pc - 1 : 0;                     // byte of the row above (i > 0 implies pc >= 1)
            const bool tieU = (i > 0) & (((nib(A, pu1) - g - hv) & 15) == 0);
            const bool tieM = inner ? (((nib(B, pu1) + sub - hv) & 15) == 0) : (hv == sub);
            const bool s0 = state == 0;
            if (depth == 0 && s0 && hv == 0) return false;           // path ends at the first zero cell
            budget -= s0 ? 1 : 0;
            // the op rule as a table over (state, op returned by the neighbour, the three ties): see decision_entry()
            const int e = dtab[state | retc | (tieL ? 16 : 0) | (tieU ? 32 : 0) | (tieM ? 64 : 0)];
            const int res = e & 3;
            const bool pushL = (e & 4) != 0, pushU = (e & 8) != 0, push = (e & 12) != 0;
            // rare exits: the stored nibble disagrees with the tracked value (e.g. window too short for the best cell);
            // a cell next to the artificial left boundary (its left / diagonal value is unknown); evaluation budget or
            // nesting limit; no tie at all (cannot happen for an exact cell)
            const bool bad = (((nib(A, pc) - hv) & 15) != 0) | ((cl <= 1) & (c0 > 0)) | (budget < 0) |
                             ((e & 16) != 0) | (push & (depth + 1 >= MAXD));
            if (bad) { failed = true; return false; }
            const bool isret = !push & (depth > 0);                  // return to the cell that asked
            const bool isres = !push & (depth == 0);                 // a path cell is resolved
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
        // after the loop: false = the task must go to the fallback list
        LSW_HD bool finish(PathOutT<OPS>& po) const {
            if (failed) return false;                                // (a started walk has evaluated at least its first cell)
            if (cl == 0 && c0 > 0 && r > 0) return false;            // ran into the artificial boundary
            po.q_pos = r;
            po.r_pos = c0 + cl;
            return true;
        }
    };

    // fill + walk for the pair of candidate records (ci0, ci1) of sequence q; ci1 may be -1.  fb(t): push a task to the fallback list.
    template <bool OPS, class Fallback>
    LSW_HD void run_pair(const TraceParams& P, const U2* prof, const uint8_t* qcodes, const uint8_t* dtab, int q, int ci0, int ci1, int64_t slot, Fallback fb) {
        Half hf[2];
        const int ci[2] = {ci0, ci1};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            hf[h].ncols = 0; hf[h].first = 0; hf[h].code0 = 0;
            if (ci[h] >= 0) {
                const TaskView v = load_cand(P, ci[h], q);
                int c0;
                window(P, v, c0, hf[h].ncols);
                hf[h].code0 = (v.ref - P.codes) + c0;
            }
        }
        const int n = hf[0].ncols > hf[1].ncols ? hf[0].ncols : hf[1].ncols;
        hf[0].first = n - hf[0].ncols;
        hf[1].first = n - hf[1].ncols;

        if (n > 0) {
            const uint32_t G2 = (uint32_t)(P.g * SC) * 0x00010001u;
            uint4* sbase = reinterpret_cast<uint4*>(P.scratch) + slot;
#pragma unroll
            for (int r = 0; r < NR; r++) M[r] = 0;
            // byte address of the read code of shared step 0 for each task (may lie before the window: masked)
            int64_t addr[2]; int lead[2]; uint32_t prev[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                addr[h] = 0; lead[h] = 0; prev[h] = 0;
                if (hf[h].ncols > 0) {
                    const int64_t a0 = hf[h].code0 - hf[h].first;
                    lead[h] = (int)(a0 & 3);
                    addr[h] = a0 - lead[h];
                    prev[h] = *reinterpret_cast<const uint32_t*>(P.codes + addr[h]);
                }
            }
#pragma unroll 1
            for (int s0 = 0; s0 < n; s0 += 4) {
                uint32_t c[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    uint32_t v = 0;
                    if (hf[h].ncols > 0) {
                        const uint32_t nxt = *reinterpret_cast<const uint32_t*>(P.codes + addr[h] + s0 + 4);
                        v = prmt(prev[h], nxt, 0x3210u + 0x1111u * (uint32_t)lead[h]);    // bytes of steps s0..s0+3
                        prev[h] = nxt;
                    }
                    if (h) v |= (v & 0x04040404u) << 2;          // stored 'other' (4) is indexed as 20 by the high half
                    int nl = hf[h].first - s0; nl = nl < 0 ? 0 : (nl > 4 ? 4 : nl);          // dead steps before the window
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
                        uint32_t d = 0, u = 0;
#pragma unroll
                        for (int r2 = 0; r2 < NR / 2; r2++) {
                            const U2 sv = prow[r2];
                            uint32_t mold = M[2 * r2];
                            uint32_t T = addmax2(d, sv.x, mold);
                            uint32_t hh = max3_2(T, u, G2);
                            u = hh - G2; M[2 * r2] = u; d = mold;
                            mold = M[2 * r2 + 1];
                            T = addmax2(d, sv.y, mold);
                            hh = max3_2(T, u, G2);
                            u = hh - G2; M[2 * r2 + 1] = u; d = mold;
                        }
                        if (!(P.debug & 2)) {
                            // H of row r: byte 1 (low task) and byte 3 (high task) of M[r]; their low nibbles are kept
                            uint4* col = sbase + (int64_t)(s0 + j) * UPAIRS * TRACE_THREADS;
                            uint32_t lo[2 * UPAIRS], hi[2 * UPAIRS];
#pragma unroll
                            for (int un = 0; un < 2 * UPAIRS; un++) {
                                uint32_t pr[4];
#pragma unroll
                                for (int k = 0; k < 4; k++) {
                                    // bytes 1 and 3 of pr[k]: nibbles of rows (ra, ra + 1) of the low / high task
                                    const int ra = 7 * un + 2 * k, rb = ra + 1;
                                    const uint32_t va = (un < UNITS && ra < NR) ? M[ra < NR ? ra : 0] : 0u;
                                    const uint32_t vb = (un < UNITS && rb < NR) ? M[rb < NR ? rb : 0] : 0u;
                                    pr[k] = (va & 0x0F000F00u) | ((vb << 4) & 0xF000F000u);
                                }
                                const uint32_t p01 = prmt(pr[0], pr[1], 0x7531u), p23 = prmt(pr[2], pr[3], 0x7531u);
                                lo[un] = prmt(p01, p23, 0x6420u);
                                hi[un] = prmt(p01, p23, 0x7531u);
                            }
#pragma unroll
                            for (int up = 0; up < UPAIRS; up++)
                                col[up * TRACE_THREADS] = make_uint4(lo[2 * up], hi[2 * up], lo[2 * up + 1], hi[2 * up + 1]);
                        }
                    }
                }
            }
        }
        if (P.debug & 3) {                                     // timing experiments only: no walk
            if (P.debug & 2) { uint32_t x = 0; for (int r = 0; r < NR; r++) x ^= M[r]; if (x == 0x12345u) fb(0); }
            return;
        }

        // the two tracebacks, one after the other; the record is read again (it is cache-hot) instead of being
        // kept in registers across the fill
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int cidx = h ? ci1 : ci0;
            if (cidx < 0) continue;
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
                g_min_cl = ncols; g_ncols_sum += ncols;
#endif
                if (w.init(P, v, c0, ncols, n - ncols, h, slot, qcodes, dtab))
                    while (w.step(P, po)) { }
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
                { int e = (ncols - g_min_cl) - v.row; e = e < -20 ? -20 : (e > 40 ? 40 : e); g_extra_hist[e + 20]++; }
#endif
                ok = w.finish(po);
            }
#if !defined(__CUDA_ARCH__) && defined(LSW_EMU_STATS)
            g_walks++; if (h == 0) g_fill_steps += n;
#endif
            if (P.debug & 4) { if (po.m == 0x7FFFFFFF) fb(v.t); continue; }      // timing experiments only: no output
            if (ok) emit(P, v.t, v, po);
            else fb(v.t);
        }
    }
};

#if defined(__CUDACC__)

template <int NR>
__host__ __device__ constexpr size_t trace_prof_bytes() {
    return sizeof(U2) * ((size_t)(TRACE_THREADS / 32) * NIDX * TraceLane<NR>::PROF_STRIDE);
}

// lower bound of `key` in the sorted 16-bit key array (keys >= 0xFF00 are padding)
__device__ __forceinline__ long long lower_bound_u16(const uint16_t* keys, long long n, uint32_t key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((uint32_t)keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Candidates arrive sorted by (sequence, window length).  Warps pick a sequence of the class (sticky, as
// in K1), load its profile, and their lanes claim consecutive PAIRS of its candidates.
// the walk is latency bound: ask the register allocator for enough resident blocks (measured on B200:
// K2 time falls almost linearly with resident warps up to the limit)
#ifndef LSW_K2_OCC_SMALL
#define LSW_K2_OCC_SMALL 6
#endif
template <int NR> struct TraceOcc {
    static constexpr int value = NR <= 24 ? LSW_K2_OCC_SMALL : NR <= 32 ? (LSW_K2_OCC_SMALL > 5 ? 5 : LSW_K2_OCC_SMALL) : 3;
};

template <int NR, bool OPS>
__global__ void __launch_bounds__(TRACE_THREADS, TraceOcc<NR>::value)
k_trace_fast(const TraceParams P) {
    using Lane = TraceLane<NR>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    U2* prof = reinterpret_cast<U2*>(smem_raw) + (size_t)warp * NIDX * Lane::PROF_STRIDE;
    // scratch is block-local: [block][window column][word][lane of the block], so a block's windows are one
    // contiguous region (a handful of pages) instead of rows strided across the whole launch
    const int64_t slot = (int64_t)blockIdx.x * P.scratch_block_words + threadIdx.x;      // in 16-byte units
    const long long ncand = (long long)*P.ncand;
    // segment of every sequence of the class in the sorted candidate list, found once per block
    long long* seg_lo = reinterpret_cast<long long*>(smem_raw + trace_prof_bytes<NR>());
    long long* seg_hi = seg_lo + 256;
    uint8_t* qsm = reinterpret_cast<uint8_t*>(seg_hi + 256) + warp * QSTRIDE;          // the warp's query codes
    uint8_t* dtab = reinterpret_cast<uint8_t*>(seg_hi + 256) + (TRACE_THREADS / 32) * QSTRIDE;
    dtab[threadIdx.x] = decision_entry((int)threadIdx.x);                              // 128 threads, 128 entries
    for (int i = threadIdx.x; i < P.n_class_seqs; i += blockDim.x) {
        const int q = P.class_seqs[i];
        seg_lo[i] = lower_bound_u16(P.cand_key, ncand, (uint32_t)q << 8);
        seg_hi[i] = lower_bound_u16(P.cand_key, ncand, ((uint32_t)q + 1) << 8);
    }
    __syncthreads();
    Lane L;
    int prefer = (int)((blockIdx.x * (TRACE_THREADS / 32) + warp) % (unsigned)P.n_class_seqs);
    for (int tried = 0; tried < P.n_class_seqs; ) {
        const int q = P.class_seqs[prefer];
        const long long lo = seg_lo[prefer];
        const long long cnt = P.fb_in ? (long long)P.fb_in_cnt[q] : seg_hi[prefer] - lo;      // second tier: what the band kernel passed on
        const uint32_t* list = (P.fb_in ? P.fb_in : P.cand) + lo;
        const long long hd0 = (long long)*((volatile unsigned long long*)&P.head[q]);
        if (hd0 >= cnt) { prefer = (prefer + 1) % P.n_class_seqs; tried++; continue; }
        tried = 0;
        __syncwarp();
        {   // profile of q (same layout as K1)
            ScoreParams SP;
            SP.qcodes = P.qcodes; SP.s_match = SC * (P.match + P.g); SP.s_mis = SC * (P.mismatch + P.g);
            ScoreLane<NR>::build_profile(SP, q, P.qlen[q], reinterpret_cast<uint32_t*>(prof), lane, 32);
            if (lane < QSTRIDE / 4)
                reinterpret_cast<uint32_t*>(qsm)[lane] = reinterpret_cast<const uint32_t*>(P.qcodes + (size_t)q * QSTRIDE)[lane];
        }
        __syncwarp();
        for (;;) {
            // one claim per warp (64 candidates), not one per lane: per-lane atomics on the queue head serialise in L2
            unsigned long long first = 0;
            if (lane == 0) first = atomicAdd(&P.head[q], 64ull);
            const long long i = (long long)__shfl_sync(0xFFFFFFFFu, first, 0) + 2 * lane;
            const bool active = i < cnt;
            if (!__any_sync(0xFFFFFFFFu, active)) break;
            if (active && P.all_fallback) {                       // scoring outside the nibble walk's range
                for (int k = 0; k < 2 && i + k < cnt; k++) {
                    const unsigned long long j = atomicAdd(P.nfallback, 1ull);
                    P.fallback[j] = (uint32_t)P.cand_rec[list[i + k]].t;
                }
            } else if (active) {
                const int c0 = (int)list[i];
                const int c1 = (i + 1 < cnt) ? (int)list[i + 1] : -1;
                L.template run_pair<OPS>(P, prof, qsm, dtab, q, c0, c1, slot, [&](int t) {
                    const unsigned long long j = atomicAdd(P.nfallback, 1ull);
                    P.fallback[j] = (uint32_t)t;
                });
            }
        }
        prefer = (prefer + 1) % P.n_class_seqs;
    }
}

template <int NR>
inline size_t trace_smem_bytes() { return trace_prof_bytes<NR>() + 2 * 256 * sizeof(long long) + (TRACE_THREADS / 32) * QSTRIDE + 128; }

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
