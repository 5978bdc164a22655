/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked or loaded by the product path.
 *
 * Plain-C restatement of the same algorithm oracle/swalign_ref.py restates in
 * Python: swalign 0.3.6 LocalAlignment.align() / Alignment.__init__ (third-party,
 * pinned /root/reference/requirements.txt:1, called at /root/reference/lamprey.py:563)
 * following SURVEY.md Appendix A.1/A.2 in the GENERAL form -- every cell carries
 * (value, op, run), gap and gap-extension penalties are separate, "no penalty to
 * start" special cases kept -- plus LAMPrey's greedy recursion
 * (/root/reference/lamprey.py:596-645) and per-read driver (lamprey.py:699-733).
 *
 * It exists so the parity tests can check millions of cells in seconds; it is
 * itself checked against the Python oracle and the upstream README vector in
 * tests/test_oracle.py.  PARITY UNPINNED by reference tests (there are none).
 *
 * Only decay == 0, prefer_gap_runs == True, local mode, no wildcard (the LAMPrey
 * parameterisation, lamprey.py:1860-1865) are restated.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score, q_pos, q_end, r_pos, r_end, matches, mismatches, n_cigar;
} swo_result;

typedef struct {
    int32_t seq;        /* index into the query list */
    int32_t start, end; /* read coordinates, end exclusive (already shifted to the whole read) */
    int32_t matches, mismatches;
    int32_t score, q_pos, q_end;
    int32_t depth;      /* recursion depth at which the hit was found */
} swo_hit;

typedef struct {
    int64_t calls, cells;
    int32_t max_depth;
} swo_stats;

static inline int up(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

/* ops: 0 ' ' / 'x' (no op), 1 'm', 2 'i', 3 'd' */
enum { OP_NONE = 0, OP_M = 1, OP_I = 2, OP_D = 3 };

static int max4(int a, int b, int c, int d) {
    int m = a > b ? a : b;
    if (c > m) m = c;
    if (d > m) m = d;
    return m;
}

/* cigar entries: (count << 2) | op with op 1 'M', 2 'I', 3 'D'. Returns 0, or -1 on
 * allocation failure; n_cigar reports the needed count even when it exceeds the cap. */
int swo_align(const uint8_t* ref, int R, const uint8_t* query, int Q,
              int match, int mismatch, int gap, int gap_ext,
              swo_result* out, uint32_t* cigar, int cigar_cap)
{
    const int ncols = R + 1, nrows = Q + 1;
    const size_t size = (size_t)nrows * (size_t)ncols;
    int32_t* val = (int32_t*)calloc(size, sizeof(int32_t));
    int32_t* run = (int32_t*)calloc(size, sizeof(int32_t));
    uint8_t* op = (uint8_t*)calloc(size, 1);
    if (!val || !run || !op) { free(val); free(run); free(op); return -1; }

    for (int r = 1; r < nrows; r++) op[(size_t)r * ncols] = OP_I;
    for (int c = 1; c < ncols; c++) op[c] = OP_D;

    int max_val = 0, max_row = 0, max_col = 0;
    for (int r = 1; r < nrows; r++) {
        const int qch = up(query[r - 1]);
        const size_t base = (size_t)r * ncols, above = base - ncols;
        for (int c = 1; c < ncols; c++) {
            const int mm_val = val[above + c - 1] + (qch == up(ref[c - 1]) ? match : mismatch);
            int ins_run = 0, del_run = 0, ins_val, del_val;

            const int up_val = val[above + c];
            if (op[above + c] == OP_I) {
                ins_run = run[above + c];
                ins_val = (up_val == 0) ? 0 : up_val + gap_ext;
            } else {
                ins_val = up_val + gap;
            }
            const int left_val = val[base + c - 1];
            if (op[base + c - 1] == OP_D) {
                del_run = run[base + c - 1];
                del_val = (left_val == 0) ? 0 : left_val + gap_ext;
            } else {
                del_val = left_val + gap;
            }

            int cell = max4(mm_val, del_val, ins_val, 0);
            int o, rn;
            if (del_run && cell == del_val)      { o = OP_D; rn = del_run + 1; }
            else if (ins_run && cell == ins_val) { o = OP_I; rn = ins_run + 1; }
            else if (cell == mm_val)             { o = OP_M; rn = 0; }
            else if (cell == del_val)            { o = OP_D; rn = 1; }
            else if (cell == ins_val)            { o = OP_I; rn = 1; }
            else                                 { cell = 0; o = OP_NONE; rn = 0; }

            if (cell >= max_val) { max_val = cell; max_row = r; max_col = c; }
            val[base + c] = cell; op[base + c] = (uint8_t)o; run[base + c] = rn;
        }
    }

    /* traceback */
    int row = max_row, col = max_col;
    int matches = 0, mismatches = 0, n_ops = 0;
    /* worst case path length is rows + cols */
    uint8_t* path = (uint8_t*)malloc((size_t)nrows + (size_t)ncols + 2);
    if (!path) { free(val); free(run); free(op); return -1; }
    for (;;) {
        const size_t i = (size_t)row * ncols + col;
        if (val[i] <= 0) break;
        const int o = op[i];
        path[n_ops++] = (uint8_t)o;
        if (o == OP_M) {
            if (up(query[row - 1]) == up(ref[col - 1])) matches++; else mismatches++;
            row--; col--;
        } else if (o == OP_I) { mismatches++; row--; }
        else if (o == OP_D)   { mismatches++; col--; }
        else break;
    }
    out->score = max_val;
    out->q_pos = row; out->r_pos = col;
    out->q_end = max_row; out->r_end = max_col;
    /* q_end / r_end as Alignment.__init__ derives them: pos + consumed length */
    {
        int ql = 0, rl = 0;
        for (int k = 0; k < n_ops; k++) {
            if (path[k] == OP_M) { ql++; rl++; }
            else if (path[k] == OP_I) ql++;
            else if (path[k] == OP_D) rl++;
        }
        out->q_end = row + ql; out->r_end = col + rl;
    }
    out->matches = matches; out->mismatches = mismatches;

    /* run-length encode in forward order */
    int n_cig = 0, cnt = 0, last = -1;
    for (int k = n_ops - 1; k >= 0; k--) {
        const int o = path[k];
        if (last >= 0 && o != last) {
            if (n_cig < cigar_cap && cigar) cigar[n_cig] = ((uint32_t)cnt << 2) | (uint32_t)last;
            n_cig++; cnt = 0;
        }
        last = o; cnt++;
    }
    if (last >= 0) {
        if (n_cig < cigar_cap && cigar) cigar[n_cig] = ((uint32_t)cnt << 2) | (uint32_t)last;
        n_cig++;
    }
    out->n_cigar = n_cig;

    free(path); free(val); free(run); free(op);
    return 0;
}

/* ---- greedy recursion (lamprey.py:596-645) ------------------------------------------- */

typedef struct {
    swo_hit* hits; int64_t cap; int64_t n;
    swo_stats* st;
    int match, mismatch, gap, gap_ext;
    double thr;
} rec_ctx;

static int rec(rec_ctx* cx, const uint8_t* seq, int len, int shift, const uint8_t* primer, int Q,
               int seq_idx, int depth)
{
    swo_result r;
    if (cx->st) {
        cx->st->calls++; cx->st->cells += (int64_t)len * Q;
        if (depth > cx->st->max_depth) cx->st->max_depth = depth;
    }
    if (swo_align(seq, len, primer, Q, cx->match, cx->mismatch, cx->gap, cx->gap_ext, &r, 0, 0)) return -1;
    const int span = r.r_end - r.r_pos;
    const int tot = r.matches + r.mismatches;
    /* identity is float(m)/(m+x), or the int 0 when nothing aligned */
    const double identity = tot > 0 ? (double)r.matches / (double)tot : 0.0;
    if (identity >= cx->thr && (double)span >= (double)Q * cx->thr) {
        if (cx->n < cx->cap) {
            swo_hit* h = &cx->hits[cx->n];
            h->seq = seq_idx; h->start = r.r_pos + shift; h->end = r.r_end + shift;
            h->matches = r.matches; h->mismatches = r.mismatches;
            h->score = r.score; h->q_pos = r.q_pos; h->q_end = r.q_end; h->depth = depth;
        }
        cx->n++;
        if (r.r_pos >= Q)
            if (rec(cx, seq, r.r_pos, shift, primer, Q, seq_idx, depth + 1)) return -1;
        if (len - r.r_end >= Q)
            if (rec(cx, seq + r.r_end, len - r.r_end, shift + r.r_end, primer, Q, seq_idx, depth + 1)) return -1;
    }
    return 0;
}

/* All hits of one read against n_seq sequences, in the reference's PRE-SORT order
 * (sequence order, pre-order within a sequence).  *n_out is the number found (may
 * exceed cap).  The caller applies the stable sort on start (lamprey.py:722). */
int swo_find_all(const uint8_t* read, int len, int n_seq, const uint8_t* concat, const int32_t* offs,
                 int match, int mismatch, int gap, int gap_ext, double thr,
                 swo_hit* hits, int64_t cap, int64_t* n_out, swo_stats* st)
{
    rec_ctx cx = { hits, cap, 0, st, match, mismatch, gap, gap_ext, thr };
    for (int s = 0; s < n_seq; s++) {
        if (rec(&cx, read, len, 0, concat + offs[s], offs[s + 1] - offs[s], s, 0)) return -1;
    }
    *n_out = cx.n;
    return 0;
}
