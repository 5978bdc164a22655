/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked or loaded by the product path.
 *
 * Plain-C restatement of the alignment behind LAMPrey's DEFAULT seeding path and its stage-2 sub-read mapping:
 *   /root/reference/lamprey.py:572-593   alignSequence_skbio:  StripedSmithWaterman(primer)(str(seq))
 *   /root/reference/lamprey.py:1651-1658 stage 2:               StripedSmithWaterman(sub-read)(target_ref_str)
 * scikit-bio's StripedSmithWaterman is a Cython wrapper (skbio/alignment/_ssw_wrapper.pyx) around Mengyao Zhao's SSW
 * library (skbio/alignment/_lib/ssw.c).  NEITHER IS PRESENT in /root/reference, in this image or on any reachable index
 * (scikit-bio is an unpinned dependency: /root/reference/requirements.txt), so what follows restates the published
 * algorithm of ssw.c from knowledge of the upstream source, function by function:
 *
 *   sw_sse2_byte / sw_sse2_word   -> sw_scalar()   the striped SIMD kernels compute the plain affine-gap recurrence
 *         H(q,i) = max(0, H(q-1,i-1) + W(read[q], ref[i]), E(q,i), F(q,i))
 *         E(q,i+1) = max(E(q,i) - gapE, H(q,i) - gapO)        (saturating at 0: unsigned arithmetic)
 *         F(q+1,i) = max(F(q,i) - gapE, H(q,i) - gapO)        (the lazy-F loop makes the striped order exact)
 *       column by column over the reference; the best score moves only when a column's maximum is STRICTLY greater than
 *       everything before it (`if (temp > max)`), that column is saved, and the read end is the SMALLEST read index of the
 *       saved column holding the maximum; `terminate` stops the scan at the first column whose maximum equals it.
 *   ssw_align                     -> sswo_align()  forward pass; reverse pass over the reversed prefixes with
 *       terminate = score1 (gives ref_begin1, read_begin1); banded_sw on the box between begin and end.
 *   banded_sw                     -> banded_sw()   restated statement by statement, including the direction codes
 *       (1 diagonal, 2/3 from E, 4/5 from F), the `temp1 <= temp2` preference for the diagonal, the band doubling
 *       `while (max < score)` and the traceback loop `while (i > 0)`.
 *   skbio wrapper defaults (gap_open 5, gap_extend 2, match 2, mismatch -3, N scores 0 against everything,
 *       zero-based inclusive begin / end, mask_length max(15, len(query) / 2))
 *
 * PARITY UNPINNED by reference tests (the reference has none).  The one published known answer this restatement is
 * checked against is the example of scikit-bio's own documentation (tests/test_oracle_ssw.py): query
 * ACTAAGGCTCTCTACCCCTCTCAGAGA on target AAAAAACTCTCTAAACTCACTAAGGCTCTCTACCCCTCTTCAGAGAAGTCGA -> score 49,
 * target 18..45, query 0..26, CIGAR 20M1D7M, suboptimal score 24 ending at target 29.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t score;                 /* optimal_alignment_score                         */
    int32_t ref_begin, ref_end;    /* target_begin, target_end_optimal (0-based, inclusive; -1 = not aligned) */
    int32_t read_begin, read_end;  /* query_begin, query_end                          */
    int32_t score2, ref_end2;      /* suboptimal_alignment_score, target_end_suboptimal */
    int32_t n_cigar;               /* runs produced (may exceed the caller's capacity) */
    int32_t matches;               /* lamprey.py:579-583: aligned columns with equal characters */
} sswo_result;

typedef struct {
    int32_t seq;
    int32_t start, end;            /* read coordinates, end exclusive: target_begin, target_end_optimal + 1 (lamprey.py:587-588) */
    int32_t matches;
    int32_t score, q_pos, q_end;   /* query_begin, query_end + 1 */
    int32_t depth;
} sswo_hit;

typedef struct { int64_t calls, cells; int32_t max_depth; int32_t pad; } sswo_stats;

/* skbio's nucleotide table: A C G T -> 0..3 (either case), everything else -> 4 ("N") */
static int nt(int c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; default: return 4; }
}

/* _build_match_matrix: 5 x 5, N row and column are 0 */
static void build_mat(int match, int mismatch, int8_t* mat) {
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++)
            mat[i * 5 + j] = (i == 4 || j == 4) ? 0 : (int8_t)(i == j ? match : mismatch);
}

typedef struct { int score, ref, read; } best_t;

/* sw_sse2_byte / sw_sse2_word, scalar.  ref_dir 0: columns 0 .. refLen-1; 1: refLen-1 .. 0.  bests[0] optimal, bests[1] the
 * "most possible 2nd best" outside the mask around the optimal end. */
static int sw_scalar(const int8_t* ref, int ref_dir, int refLen, const int8_t* read, int readLen, int gapO, int gapE,
                     const int8_t* mat, int terminate, int maskLen, best_t* bests)
{
    int max = 0, end_read = readLen - 1, end_ref = -1;
    int* Hp = (int*)calloc((size_t)(readLen > 0 ? readLen : 1), sizeof(int));
    int* Hc = (int*)calloc((size_t)(readLen > 0 ? readLen : 1), sizeof(int));
    int* E = (int*)calloc((size_t)(readLen > 0 ? readLen : 1), sizeof(int));
    int* Hmax = (int*)calloc((size_t)(readLen > 0 ? readLen : 1), sizeof(int));
    int* maxColumn = (int*)calloc((size_t)(refLen > 0 ? refLen : 1), sizeof(int));
    if (!Hp || !Hc || !E || !Hmax || !maxColumn) { free(Hp); free(Hc); free(E); free(Hmax); free(maxColumn); return -1; }
    int begin = 0, end = refLen, step = 1;
    if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }
    for (int i = begin; i != end; i += step) {
        int F = 0, colmax = 0;
        for (int q = 0; q < readLen; q++) {
            int h = (q > 0 ? Hp[q - 1] : 0) + mat[ref[i] * 5 + read[q]];
            if (h < 0) h = 0;                         /* saturating unsigned arithmetic */
            if (E[q] > h) h = E[q];
            if (F > h) h = F;
            if (h > colmax) colmax = h;
            Hc[q] = h;
            int hg = h - gapO; if (hg < 0) hg = 0;
            int e = E[q] - gapE; if (e < 0) e = 0;
            E[q] = e > hg ? e : hg;
            int f = F - gapE; if (f < 0) f = 0;
            F = f > hg ? f : hg;
        }
        if (colmax > max) {                           /* `if (LIKELY(temp > max))`: strictly greater */
            max = colmax;
            end_ref = i;
            memcpy(Hmax, Hc, (size_t)readLen * sizeof(int));
        }
        maxColumn[i] = colmax;
        { int* t = Hp; Hp = Hc; Hc = t; }
        if (colmax == terminate) break;
    }
    /* Trace the alignment ending position on read: the smallest index holding the maximum */
    for (int q = 0; q < readLen; q++)
        if (Hmax[q] == max) { if (q < end_read) end_read = q; }
    bests[0].score = max; bests[0].ref = end_ref; bests[0].read = end_read;
    bests[1].score = 0; bests[1].ref = 0; bests[1].read = 0;
    {   /* Find the most possible 2nd best alignment. */
        int edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
        for (int i = 0; i < edge; i++)
            if (maxColumn[i] > bests[1].score) { bests[1].score = maxColumn[i]; bests[1].ref = i; }
        edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
        for (int i = edge + 1; i < refLen; i++)
            if (maxColumn[i] > bests[1].score) { bests[1].score = maxColumn[i]; bests[1].ref = i; }
    }
    free(Hp); free(Hc); free(E); free(Hmax); free(maxColumn);
    return 0;
}

#define set_u(u, w, i, j) { int x_ = (i) - (w); x_ = x_ > 0 ? x_ : 0; (u) = (j) - x_ + 1; }
#define set_d(u, w, i, j, p) { int x_ = (i) - (w); x_ = x_ > 0 ? x_ : 0; x_ = (j) - x_; (u) = x_ * 3 + (p); }

/* banded_sw.  cigar runs are (count << 2) | op, op 1 'M', 2 'I', 3 'D', first run first.  Returns the number of runs, -1 on
 * allocation failure, -2 if the band cannot reach `score` any more (the box is not the box of an alignment with that score:
 * upstream would double the band for ever). */
static int banded_sw(const int8_t* ref, const int8_t* read, int refLen, int readLen, int score, int gapO, int gapE,
                     int band_width, const int8_t* mat, uint32_t* cigar, int cap, int* n_out)
{
    int i, j, e, f, temp1, temp2, l, max = 0;
    int width = 0, width_d = 0;
    int *h_b = 0, *e_b = 0, *h_c = 0;
    int8_t *direction = 0, *direction_line;
    for (;;) {
        width = band_width * 2 + 3; width_d = band_width * 2 + 1;
        free(h_b); free(e_b); free(h_c); free(direction);
        h_b = (int*)calloc((size_t)width + 1, sizeof(int)); e_b = (int*)calloc((size_t)width + 1, sizeof(int));
        h_c = (int*)calloc((size_t)width + 1, sizeof(int));
        direction = (int8_t*)calloc((size_t)width_d * (size_t)readLen * 3 + 3, 1);
        if (!h_b || !e_b || !h_c || !direction) { free(h_b); free(e_b); free(h_c); free(direction); return -1; }
        for (j = 1; j < width - 1; j++) h_b[j] = 0;
        for (i = 0; i < readLen; i++) {
            int beg = 0, end = refLen - 1, u = 0, edge;
            j = i - band_width; beg = beg > j ? beg : j;        /* band start */
            j = i + band_width; end = end < j ? end : j;        /* band end */
            edge = end + 1 < width - 1 ? end + 1 : width - 2;
            f = h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
            direction_line = direction + (size_t)width_d * i * 3;
            for (j = beg; j <= end; j++) {
                int b, e1, f1, d, de, df, dh;
                set_u(u, band_width, i, j); set_u(e, band_width, i - 1, j);
                set_u(b, band_width, i, j - 1); set_u(d, band_width, i - 1, j - 1);
                set_d(de, band_width, i, j, 0);
                set_d(df, band_width, i, j, 1);
                set_d(dh, band_width, i, j, 2);

                temp1 = i == 0 ? -gapO : h_b[e] - gapO;
                temp2 = i == 0 ? -gapE : e_b[e] - gapE;
                e_b[u] = temp1 > temp2 ? temp1 : temp2;
                direction_line[de] = temp1 > temp2 ? 3 : 2;

                temp1 = h_c[b] - gapO;
                temp2 = f - gapE;
                f = temp1 > temp2 ? temp1 : temp2;
                direction_line[df] = temp1 > temp2 ? 5 : 4;

                e1 = e_b[u] > 0 ? e_b[u] : 0;
                f1 = f > 0 ? f : 0;
                temp1 = e1 > f1 ? e1 : f1;
                temp2 = h_b[d] + mat[ref[j] * 5 + read[i]];
                h_c[u] = temp1 > temp2 ? temp1 : temp2;

                if (h_c[u] > max) max = h_c[u];

                if (temp1 <= temp2) direction_line[dh] = 1;
                else direction_line[dh] = e1 > f1 ? direction_line[de] : direction_line[df];
            }
            for (j = 1; j <= u; j++) h_b[j] = h_c[j];
        }
        if (max >= score) break;
        if (band_width > refLen + readLen) { free(h_b); free(e_b); free(h_c); free(direction); return -2; }
        band_width *= 2;
    }

    /* trace back */
    uint32_t* c = (uint32_t*)malloc(((size_t)refLen + readLen + 4) * sizeof(uint32_t));
    if (!c) { free(h_b); free(e_b); free(h_c); free(direction); return -1; }
    i = readLen - 1;
    j = refLen - 1;
    e = 0;      /* Count the number of M, D or I. */
    l = 0;      /* record length of current cigar */
    int op = 1, prev_op = 1;
    temp2 = 2;  /* h */
    direction_line = direction + (size_t)width_d * i * 3;
    int bad = 0;
    while (i > 0) {
        set_d(temp1, band_width, i, j, temp2);
        if (j < 0 || temp1 < 0 || temp1 >= width_d * 3) { bad = 1; break; }      /* upstream would read out of bounds */
        switch (direction_line[temp1]) {
            case 1: --i; --j; temp2 = 2; direction_line -= width_d * 3; op = 1; break;
            case 2: --i; temp2 = 0; direction_line -= width_d * 3; op = 2; break;      /* e */
            case 3: --i; temp2 = 2; direction_line -= width_d * 3; op = 2; break;
            case 4: --j; temp2 = 1; op = 3; break;
            case 5: --j; temp2 = 2; op = 3; break;
            default: bad = 1; break;                                                    /* "Trace back error" */
        }
        if (bad) break;
        if (op == prev_op) ++e;
        else {
            ++l;
            c[l - 1] = ((uint32_t)e << 2) | (uint32_t)prev_op;
            prev_op = op;
            e = 1;
        }
    }
    if (!bad) {
        if (op == 1) { ++l; c[l - 1] = ((uint32_t)(e + 1) << 2) | 1u; }
        else { l += 2; c[l - 2] = ((uint32_t)e << 2) | (uint32_t)op; c[l - 1] = (1u << 2) | 1u; }
    }
    /* reverse cigar */
    int n = bad ? 0 : l;
    for (int k = 0; k < n; k++) if (k < cap) cigar[k] = c[n - 1 - k];
    *n_out = n;
    free(c); free(h_b); free(e_b); free(h_c); free(direction);
    return bad ? -2 : n;
}

/* ssw_align with flag 1 (skbio's default bit flag: begin positions and cigar) on raw ASCII.  Returns 0, -1 out of memory,
 * -2 the banded pass could not reproduce the score (see banded_sw; out->n_cigar = 0 then). */
int sswo_align(const uint8_t* target, int refLen, const uint8_t* query, int readLen, int match, int mismatch, int gapO,
               int gapE, int maskLen, sswo_result* out, uint32_t* cigar, int cigar_cap)
{
    int8_t mat[25];
    build_mat(match, mismatch, mat);
    int8_t* ref = (int8_t*)malloc((size_t)refLen + 1);
    int8_t* read = (int8_t*)malloc((size_t)readLen + 1);
    int8_t* rev = (int8_t*)malloc((size_t)readLen + 1);
    if (!ref || !read || !rev) { free(ref); free(read); free(rev); return -1; }
    for (int i = 0; i < refLen; i++) ref[i] = (int8_t)nt(target[i]);
    for (int i = 0; i < readLen; i++) read[i] = (int8_t)nt(query[i]);
    memset(out, 0, sizeof *out);
    out->ref_begin = -1; out->read_begin = -1;
    if (maskLen < 15) maskLen = 15;                                    /* ssw_align: "When maskLen < 15, the function ... 15" */
    best_t bests[2], rbests[2];
    int rc = 0;
    if (sw_scalar(ref, 0, refLen, read, readLen, gapO, gapE, mat, -1, maskLen, bests)) { rc = -1; goto done; }
    out->score = bests[0].score; out->ref_end = bests[0].ref; out->read_end = bests[0].read;
    out->score2 = bests[1].score; out->ref_end2 = bests[1].ref;
    if (refLen == 0 || readLen == 0 || bests[0].ref < 0) { out->ref_end = -1; out->read_end = readLen - 1; goto done; }   /* nothing aligned */
    /* Find the beginning position of the best alignment: the same scan over the reversed prefixes */
    for (int k = 0; k <= out->read_end; k++) rev[k] = read[out->read_end - k];
    if (sw_scalar(ref, 1, out->ref_end + 1, rev, out->read_end + 1, gapO, gapE, mat, out->score, maskLen, rbests)) { rc = -1; goto done; }
    out->ref_begin = rbests[0].ref;
    out->read_begin = out->read_end - rbests[0].read;
    {
        const int rl = out->ref_end - out->ref_begin + 1, ql = out->read_end - out->read_begin + 1;
        int band = abs(rl - ql) + 1, n = 0;
        const int r = banded_sw(ref + out->ref_begin, read + out->read_begin, rl, ql, out->score, gapO, gapE, band, mat, cigar,
                                cigar_cap, &n);
        if (r == -1) { rc = -1; goto done; }
        if (r == -2) { rc = -2; out->n_cigar = 0; goto done; }
        out->n_cigar = n;
        /* lamprey.py:579-583 on skbio's aligned_query_sequence / aligned_target_sequence: only 'M' columns can compare equal */
        int ti = out->ref_begin, qi = out->read_begin, m = 0;
        for (int k = 0; k < n && k < cigar_cap; k++) {
            const int len = (int)(cigar[k] >> 2), op = (int)(cigar[k] & 3u);
            if (op == 1) { for (int t = 0; t < len; t++) { if (target[ti + t] == query[qi + t]) m++; } ti += len; qi += len; }
            else if (op == 2) qi += len;
            else ti += len;
        }
        out->matches = m;
    }
done:
    free(ref); free(read); free(rev);
    return rc;
}

/* ---- greedy recursion with the skbio branch (lamprey.py:596-645 with args.swalign False) ------------------------------- */
typedef struct {
    sswo_hit* hits; int64_t cap; int64_t n;
    sswo_stats* st;
    int match, mismatch, gapO, gapE;
    double thr;
    uint32_t cig[512];      /* CIGAR scratch of the call in progress: nothing reads it after sswo_align() returns, and a frame without
                               it is ~100 bytes -- exact tandem repeats of a short primer recurse once per copy (thousands deep on a
                               40 kb read; the reference's Python would have hit its recursion limit long before) */
} rec_ctx;

static int rec(rec_ctx* cx, const uint8_t* seq, int len, int shift, const uint8_t* primer, int Q, int seq_idx, int depth)
{
    sswo_result r;
    uint32_t* cig = cx->cig;
    if (cx->st) {
        cx->st->calls++; cx->st->cells += (int64_t)len * Q;
        if (depth > cx->st->max_depth) cx->st->max_depth = depth;
    }
    const int mask = Q / 2 > 15 ? Q / 2 : 15;
    const int rc = sswo_align(seq, len, primer, Q, cx->match, cx->mismatch, cx->gapO, cx->gapE, mask, &r, cig, 512);
    if (rc == -1) return -1;
    if (rc == -2 || r.ref_begin < 0) return 0;           /* no usable alignment: the identity of an empty alignment is 0 */
    const double identity = (double)r.matches / (double)Q;                /* float(matches)/float(len(primer)) */
    const int start = r.ref_begin, end = r.ref_end + 1;
    if (identity >= cx->thr && (double)(end - start) >= (double)Q * cx->thr) {
        if (cx->n < cx->cap) {
            sswo_hit* h = &cx->hits[cx->n];
            h->seq = seq_idx; h->start = start + shift; h->end = end + shift; h->matches = r.matches;
            h->score = r.score; h->q_pos = r.read_begin; h->q_end = r.read_end + 1; h->depth = depth;
        }
        cx->n++;
        if (start >= Q)
            if (rec(cx, seq, start, shift, primer, Q, seq_idx, depth + 1)) return -1;
        if (len - end >= Q)
            if (rec(cx, seq + end, len - end, shift + end, primer, Q, seq_idx, depth + 1)) return -1;
    }
    return 0;
}

int sswo_find_all(const uint8_t* read, int len, int n_seq, const uint8_t* concat, const int32_t* offs, int match, int mismatch,
                  int gapO, int gapE, double thr, sswo_hit* hits, int64_t cap, int64_t* n_out, sswo_stats* st)
{
    rec_ctx cx = { hits, cap, 0, st, match, mismatch, gapO, gapE, thr, {0} };
    for (int s = 0; s < n_seq; s++)
        if (rec(&cx, read, len, 0, concat + offs[s], offs[s + 1] - offs[s], s, 0)) return -1;
    *n_out = cx.n;
    return 0;
}
