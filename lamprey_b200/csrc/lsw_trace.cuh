// lsw_trace.cuh -- K2: bounded-window traceback of swalign.LocalAlignment.align()
//
// Replaces the traceback loop, _reduce_cigar and Alignment.__init__ of the third-party
// swalign module behind /root/reference/lamprey.py:563-566 (semantics: SURVEY.md Appendix
// A.1, A.2, A.4), and -- in find_all mode -- the hit test and the left/right split of
// findPrimerAlignments (/root/reference/lamprey.py:612-640).
//
// K1 (lsw_score.cuh) has already produced the exact (score, max_row, max_col).  This kernel
// recomputes the matrix on the last W columns ending at max_col, this time with swalign's
// op rule, stores a 2-bit op per cell and walks back from (max_row, max_col).
//
// Why a window is exact (gap == gap_ext = -g, match = M, query length Q):
//   * every prefix of a traceback path has a positive score, so it contains fewer than
//     M*Q/g read-only ('D') moves: span < Q + M*Q/g =: S;
//   * a cell's VALUE therefore only depends on the S columns to its left;
//   * a cell's OP additionally depends on the op of its left/up neighbour only when their
//     values tie, and along such a chain values strictly increase by g, so the chain reaches
//     fewer than M*Q/g =: C columns further left.
//   The walk touches columns > max_col - S; their ops need values back to max_col - S - C,
//   and those values are exact once the window started S columns earlier:  W = 2S + C + 4.
//   (For LAMPrey's +2/-1/-1 this is the 8Q of SURVEY.md Appendix A.4.)  A window that starts
//   at the slice start is the true boundary and needs no argument.
//
// swalign's op rule for a positive cell (prefer_gap_runs=True):
//   'd' if left is a positive 'd' cell and H == L - g, else 'i' if up is a positive 'i' cell
//   and H == U - g, else 'm' if H == D + s, else 'd' if H == L - g, else 'i'.
#pragma once
#include "lsw_common.cuh"
#include "../../include/lsw.h"

namespace lsw {

LSW_HD int trace_window(int Q, int match, int g) {
    const int dmax = (match * Q) / g;          // >= number of 'D' moves on any positive-prefix path
    const int span = Q + dmax;
    return 2 * span + dmax + 4;
}

enum { OPC_NONE = 0, OPC_M = 1, OPC_I = 2, OPC_D = 3 };

template <int NR>
LSW_HD void trace_task(const TraceParams& P, const int t, const int64_t slot) {
    static_assert(NR % 16 == 0 && NR <= 64, "one 32-bit word of 2-bit ops per 16 rows");
    constexpr int WORDS = NR / 16;
    const int rd = P.t.read[t];
    const int a = P.t.a[t], b = P.t.b[t];
    const int q = P.t.seq[t];
    const int2 rs = P.res[t];
    const int score = rs.x & 0xFFFF;
    const int row = rs.x >> 16;       // max_row (1-based DP row)  == q_end
    const int col = rs.y;             // max_col (1-based DP col)  == r_end, slice relative
    const int Q = P.qlen[q];
    const uint8_t* ref = P.codes + P.roff[rd] + a;
    const uint8_t* qc = P.qcodes + (size_t)q * QSTRIDE;

    int q_pos = row, r_pos = col, m = 0, x = 0, nrun = 0;
    int c0 = 0;

    if (score > 0) {
        const int W = trace_window(Q, P.match, P.g);
        c0 = col > W ? col - W : 0;
        const int ncols = col - c0;

        // query codes, 4 per word; rows >= Q hold 0xFF and never match
        uint32_t qw[NR / 4];
#pragma unroll
        for (int i = 0; i < NR / 4; i++) qw[i] = reinterpret_cast<const uint32_t*>(qc)[i];

        int H[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) H[r] = 0;
        unsigned long long dprev = 0;       // bit r: cell (r, previous column) is a positive 'd' cell

        for (int c = 1; c <= ncols; c++) {
            const int bc = ref[c0 + c - 1];
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
        int r = row, c = ncols, last = 0;
        while (r > 0 && c > 0) {
            const uint32_t wv = P.scratch[((int64_t)(c - 1) * WORDS + ((r - 1) >> 4)) * P.scratch_stride + slot];
            const int op = (int)((wv >> (2 * ((r - 1) & 15))) & 3u);
            if (op == OPC_NONE) break;
            if (op != last) { nrun++; last = op; }
            if (op == OPC_M) {
                if (qc[r - 1] == ref[c0 + c - 1]) m++; else x++;
                r--; c--;
            } else if (op == OPC_I) { x++; r--; }
            else { x++; c--; }
        }
        if (c == 0 && c0 > 0 && r > 0) atomicExch_or_store(P.errflag, 1);   // window bound violated: never expected
        q_pos = r;
        r_pos = c0 + c;
    }

    if (P.mode == 1) {
        lsw_result* out = reinterpret_cast<lsw_result*>(P.results) + P.t.orig[t];
        out->score = score;
        out->q_pos = q_pos; out->q_end = row;
        out->r_pos = r_pos; out->r_end = col;
        out->matches = m; out->mismatches = x;
        out->n_cigar = nrun;
        long long off = 0;
        if (P.cigar && nrun > 0) {
            off = (long long)atomic_add_ull(P.ncigar, (unsigned long long)nrun);
            if (off + nrun <= P.cigar_cap) {
                // second walk: runs are met last-to-first
                const int ncols = col - c0;
                int r = row, c = ncols, last = 0, cnt = 0, k = nrun;
                while (r > 0 && c > 0) {
                    const uint32_t wv = P.scratch[((int64_t)(c - 1) * WORDS + ((r - 1) >> 4)) * P.scratch_stride + slot];
                    const int op = (int)((wv >> (2 * ((r - 1) & 15))) & 3u);
                    if (op == OPC_NONE) break;
                    if (op != last) {
                        if (last) P.cigar[off + (--k)] = ((uint32_t)cnt << 2) | (uint32_t)last;
                        last = op; cnt = 0;
                    }
                    cnt++;
                    if (op == OPC_M) { r--; c--; } else if (op == OPC_I) r--; else c--;
                }
                if (last) P.cigar[off + (--k)] = ((uint32_t)cnt << 2) | (uint32_t)last;
            }
        }
        out->cigar_off = off;
        return;
    }

    // ---- find_all mode: hit test and split (lamprey.py:612-640) ---------------------------
    const int tot = m + x;
    const int span = col - r_pos;
    // identity is float(m)/(m+x) (Python double division), or the int 0 when nothing aligned
    const double identity = tot > 0 ? (double)m / (double)tot : 0.0;
    if (identity >= P.thr && (double)span >= (double)Q * P.thr) {
        const unsigned long long hi = atomic_add_ull(P.nhits, 1ull);
        if ((long long)hi < P.hits_cap) {
            lsw_hit* h = reinterpret_cast<lsw_hit*>(P.hits) + hi;
            h->read = rd; h->start = a + r_pos; h->end = a + col;
            h->query = (int16_t)q; h->depth = (int16_t)P.round;
            h->matches = (int16_t)m; h->mismatches = (int16_t)x;
            h->score = (int16_t)score; h->q_pos = (int16_t)q_pos; h->q_end = (int16_t)row;
            h->reserved = 0; h->reserved2 = 0;
        }
        if (r_pos >= Q) {                       // len(left_seq) >= len(primer)
            const int64_t j = 2 * (int64_t)t;
            P.child.read[j] = rd; P.child.a[j] = a; P.child.b[j] = a + r_pos; P.child.seq[j] = (int16_t)q;
            P.child_flag[j] = 1;
        }
        if (b - (a + col) >= Q) {               // len(right_seq) >= len(primer)
            const int64_t j = 2 * (int64_t)t + 1;
            P.child.read[j] = rd; P.child.a[j] = a + col; P.child.b[j] = b; P.child.seq[j] = (int16_t)q;
            P.child_flag[j] = 1;
        }
    }
}

#if defined(__CUDACC__)
constexpr int TRACE_THREADS = 128;

template <int NR>
__global__ void __launch_bounds__(TRACE_THREADS)
k_trace(const TraceParams P) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n = *P.ncand;
    for (unsigned long long i = (unsigned long long)slot; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        trace_task<NR>(P, (int)P.cand[i], slot);
}
#endif

}  // namespace lsw
