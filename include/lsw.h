/* lsw.h -- C ABI of the B200 batched Smith-Waterman seeding engine (liblsw.so).
 *
 * This is the drop-in boundary for the hot path of jackwadden/lamprey's seed stage.
 * The reference reaches that path through a PYTHON MODULE boundary, not an FFI:
 *
 *   /root/reference/lamprey.py:14-16     sys.path.insert(0, './lib'); import swalign
 *   /root/reference/lamprey.py:1860-1865 swalign.NucleotideScoringMatrix(2, -1); swalign.LocalAlignment(scoring)
 *   /root/reference/lamprey.py:563       aligner.align(str(seq), primer)          (ref = read slice, query = primer)
 *   /root/reference/lamprey.py:596-645   findPrimerAlignments   (greedy left/right recursion around the best hit)
 *   /root/reference/lamprey.py:699-733   findAllPrimerAlignments (T, Tc, every schema entry fwd + rc; stable sort)
 * and, for the rows next to it:
 *   /root/reference/lamprey.py:17, 572-593   from skbio.alignment import StripedSmithWaterman; alignSequence_skbio (default branch)
 *   /root/reference/lamprey.py:859-998       removeOverlappingPrimerAlignments, extractAmpliconAroundTarget
 *   /root/reference/lamprey.py:1566-1669     processLamplicon stage 1 (sub-reads) and stage 2 (sub-read vs target reference)
 *   /root/reference/lamprey.py:2182-2207     reads dealt over worker processes, results collected from a queue
 *
 * so the entry points below are what a replacement `swalign` module (lib/swalign.py in this
 * repo; binding stub in INTEGRATION.md) binds through ctypes.  Plain pointers and sizes only;
 * the caller owns every host buffer, the library owns device memory inside the context.
 * All functions return 0 on success and a negative LSW_E_* code on failure; the message is
 * available from lsw_last_error().  There is NO CPU fallback: without a CUDA device
 * lsw_create() fails with LSW_E_CUDA.
 *
 * A context is single-threaded (the reference creates one aligner per forked worker,
 * lamprey.py:1886, and calls it synchronously); CUDA is initialised in lsw_create(), never
 * at library load, so a forked child may create its own context.
 */
#ifndef LSW_H
#define LSW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSW_OK            0
#define LSW_E_CUDA       -1   /* no device / CUDA runtime or launch failure            */
#define LSW_E_ARG        -2   /* bad argument (NULL, negative size, out-of-range index) */
#define LSW_E_UNSUPPORTED -3  /* scoring / query the kernels do not implement           */
#define LSW_E_OVERFLOW   -4   /* caller buffer too small; *n_out / *cigar_used = needed */
#define LSW_E_STATE      -5   /* call order (e.g. align before set_queries)             */
#define LSW_E_INTERNAL   -6   /* an internal invariant failed (reported, never silent)  */

#define LSW_MAX_ORDER 16       /* names in a fwd_primer_order / rev_primer_order line (lsw_chain_subreads) */
#define LSW_MAX_QUERY_LEN 63  /* rows a query may have (kernels keep all rows in registers) */

typedef struct lsw_ctx lsw_ctx;

/* One local alignment request: schema sequence `query` against read[start:end).
 * Replaces one swalign.LocalAlignment.align(ref, query) call (lamprey.py:563). */
typedef struct {
    int32_t read;      /* index into the uploaded read batch        */
    int32_t start;     /* slice start on the read (0-based)         */
    int32_t end;       /* slice end, exclusive                      */
    int32_t query;     /* index into the sequences of lsw_set_queries */
} lsw_task;

/* Fields of swalign.Alignment that LAMPrey consumes (lamprey.py:564-566, 672-694).
 * Positions are relative to the SLICE, 0-based, ends exclusive, exactly as swalign reports. */
typedef struct {
    int32_t score;
    int32_t q_pos, q_end;
    int32_t r_pos, r_end;
    int32_t matches, mismatches;   /* identity = matches / (matches + mismatches), 0 if both 0 */
    int32_t n_cigar;               /* number of (count, op) runs                             */
    int64_t cigar_off;             /* offset of the first run in the caller's cigar buffer     */
} lsw_result;

/* One seed hit of findPrimerAlignments (lamprey.py:138-161 Alignment + the swalign fields behind it).
 * start/end are READ coordinates (already shifted, lamprey.py:637-640). */
typedef struct {
    int32_t read;
    int32_t start, end;
    int16_t query;                 /* index into the sequences of lsw_set_queries */
    int16_t depth;                 /* recursion depth at which the hit was found  */
    int16_t matches, mismatches;
    int16_t score;
    int16_t q_pos, q_end;
    int16_t reserved;
    int32_t reserved2;
} lsw_hit;                         /* 32 bytes */

/* The same hit in 16 bytes (half the device -> host traffic of a batch; ~128 hits per 2 kb concatemer read).
 * end = start + span; score = match*matches + mismatch*mm - |gap|*(I + D) with
 *   a = span + (q_end - q_pos) - matches - mismatches   (aligned columns),   mm = a - matches,
 *   D = span - a,   I = (q_end - q_pos) - a. */
typedef struct {
    uint32_t read;
    uint32_t start;
    uint8_t  span;                 /* end - start: a path spans < Q + match*Q/|gap| <= 190 read columns */
    uint8_t  query;
    uint16_t depth;                /* saturates at 65535 */
    uint8_t  matches;              /* bits 0-6; bit 7 = the `reserved` survivor flag of lsw_set_overlap */
    uint8_t  mismatches;
    uint8_t  q_pos, q_end;
} lsw_hit16;                       /* 16 bytes */

/* One candidate sub-read of stage 1 of processLamplicon (lamprey.py:1566-1589): the amplicon extractAmpliconAroundTarget
 * grows around a surviving target hit ('T' or 'Tc'), in read coordinates as the reference slices them (seq[start:end]). */
typedef struct {
    int32_t read;
    int32_t start, end;            /* end: amplicon_end, or len(read) - 1 where the reference got -1 */
    int32_t target_idx;            /* n-th target of the read: the subread id is "<lamp_idx>_<target_idx>" */
    int32_t direction;             /* 0 'fwd' (target 'T'), 1 'rev' (target 'Tc') */
    int32_t hit;                   /* the target hit: index within the read's hits */
} lsw_subread;

/* One stage-2 alignment (lamprey.py:1637-1658): the QUERY is read[start:end) of the uploaded batch -- a candidate sub-read --
 * reverse-complemented first if revcomp != 0 ('rev' targets, lamprey.py:1642-1643). */
typedef struct {
    int32_t read;
    int32_t start, end;
    int32_t revcomp;
} lsw_pair;

typedef struct {
    int64_t tasks;        /* align() invocations the reference would have made            */
    int64_t cells;        /* sum over those of len(slice) * len(query)  (the CUPS unit)   */
    int64_t top_tasks;    /* round-0 (whole read) invocations                            */
    int64_t top_cells;
    int64_t traced;       /* tasks that needed a traceback (score could pass threshold)   */
    int64_t hits;
    int32_t rounds;       /* recursion depth reached + 1                                  */
    int32_t reserved;
    float   ms_score;     /* device time in the score kernels (CUDA events, own stream)   */
    float   ms_trace;     /* device time in the traceback kernels                         */
    float   ms_other;     /* compaction, sorting, bookkeeping kernels                     */
    float   ms_total;     /* device time of the whole call                                */
    int64_t score_launches, trace_launches, other_launches;
    int64_t steps;        /* lane-steps (columns incl. dead padding) the score kernels ran */
    int64_t fallbacks;    /* traced tasks the short-window traceback could not prove exact and that
                             were redone by the full-window kernel                              */
    int64_t computed_cells; /* cells the score kernels actually filled: child tasks reuse the
                             round-0 matrix of their read (lsw_reuse.cuh), so this is < cells     */
    int64_t device_bytes; /* device memory the context holds after this call (high-water: buffers only grow) */
    int64_t reuse_dropped;/* 1 if the block bests of lsw_reuse.cuh could not be allocated and every round aligned
                             whole slices instead (same results, more cells); 0 otherwise               */
    float   ms_score_round0; /* part of ms_score spent on whole reads (round 0); the rest ran on child segments */
    float   reserved3;
    int64_t computed_cells_round0;
} lsw_counters;

/* cigar runs are (count << 2) | op with op 1 = 'M', 2 = 'I', 3 = 'D' */

int  lsw_create(int device, lsw_ctx** out);
void lsw_destroy(lsw_ctx* ctx);
const char* lsw_last_error(const lsw_ctx* ctx);      /* ctx may be NULL: error of the last failed lsw_create */

/* Use an existing CUDA stream (cudaStream_t passed as void*; NULL = the context's own stream). */
int lsw_set_stream(lsw_ctx* ctx, void* cuda_stream);

/* swalign.NucleotideScoringMatrix(match, mismatch) + LocalAlignment(gap_penalty, gap_extension_penalty).
 * Supported: match > 0, mismatch < 0, gap == gap_ext < 0, and match*len(query) + |gap| <= 127 for every query.
 * Scoring and queries may be set in either order; lsw_find_all / lsw_align_tasks return LSW_E_UNSUPPORTED when the
 * pair in force does not fit. */
int lsw_set_scoring(lsw_ctx* ctx, int match, int mismatch, int gap, int gap_ext);

/* Switches the context to LAMPrey's DEFAULT seeding path (no --swalign): scikit-bio's StripedSmithWaterman(primer)(read slice)
 * (/root/reference/lamprey.py:572-593, 609-610).  skbio's defaults are match 2, mismatch -3, gap_open 5, gap_extend 2 (penalties
 * positive, as skbio takes them); a non-ACGT read base scores 0 against everything.  In this mode
 *   lsw_align_tasks  returns ssw_align's fields: score, r_pos / r_end = target_begin / target_end_optimal + 1, q_pos / q_end =
 *                    query_begin / query_end + 1 (r_pos = -1 if nothing aligned), matches = aligned columns with equal bases
 *                    (lamprey.py:579-583), mismatches = the other CIGAR columns, CIGAR as banded_sw produces it;
 *   lsw_find_all     applies identity = matches / len(primer) >= thr and (end - start) >= len(primer) * thr, and recurses.
 * lsw_set_scoring() switches back to the swalign path. */
int lsw_set_ssw(lsw_ctx* ctx, int match, int mismatch, int gap_open, int gap_extend);

/* The schema sequences (already in search order; reverse complements supplied by the caller).
 * ASCII, upper-cased by the library; must be ACGT only and 1..LSW_MAX_QUERY_LEN long. */
int lsw_set_queries(lsw_ctx* ctx, int n, const uint8_t* concat, const int32_t* offs /* n+1 */);

/* Host -> device copy of a read batch (ASCII, any bytes; non-ACGT never matches) and device-side
 * re-coding.  offs has n+1 entries.  Replaces the per-read str(seq) handed to align(). */
int lsw_upload_reads(lsw_ctx* ctx, int64_t n, const uint8_t* ascii, const int64_t* offs);

/* Batched align(): n_tasks independent alignments, results in task order.  cigar_buf may be NULL
 * (no CIGARs wanted).  *cigar_used is the number of runs produced; LSW_E_OVERFLOW if > cigar_cap. */
int lsw_align_tasks(lsw_ctx* ctx, int64_t n_tasks, const lsw_task* tasks, lsw_result* out,
                    uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);

/* findAllPrimerAlignments for every uploaded read: all queries, greedy recursion on the device.
 * Hits are returned sorted by (read, start, query) -- the order lamprey.py:722's stable sort
 * produces.  Device-resident variant: results stay on the device until lsw_fetch_hits(). */
int lsw_set_read_base(lsw_ctx* ctx, int64_t base);     /* added to lsw_hit.read: a caller that splits one FASTQ
                                                           batch over several contexts / GPUs gets global read ids */
/* removeOverlappingPrimerAlignments (lamprey.py:859-899, called with 4 at lamprey.py:1536) on the device: with
 * allowed_overlap >= 0 every hit lsw_find_all returns has reserved = 1 if it survives the pruning of its read's
 * sorted hit list, 0 otherwise.  Negative (default) = off. */
int lsw_set_overlap(lsw_ctx* ctx, int allowed_overlap);
/* Stage 1 of processLamplicon for every read of the last lsw_find_all (which must have run with lsw_set_overlap >= 0; LAMPrey
 * uses 4): extractAmpliconAroundTarget around every surviving target hit.  fwd_order / rev_order: the query index of every name
 * of the schema's fwd_primer_order / rev_primer_order lines (-1 for a name that is not a schema sequence); t_query / tc_query:
 * the query indices of 'T' and 'Tc'.  Sub-reads come back ordered by (read, target_idx). */
int lsw_chain_subreads(lsw_ctx* ctx, int n_fwd, const int32_t* fwd_order, int n_rev, const int32_t* rev_order, int t_query,
                       int tc_query, lsw_subread* out, int64_t cap, int64_t* n_out);
/* Stage 2 of processLamplicon: StripedSmithWaterman(sub-read)(target_ref_str) for n sub-reads of the uploaded batch against ONE
 * target reference (ASCII, target_len bases), with the scoring of lsw_set_ssw (skbio defaults if it was never called).  The
 * query may be of any length.  Results as lsw_align_tasks returns them in SSW mode: r_pos / r_end on the TARGET (target_begin,
 * target_end_optimal + 1; r_pos = -1: nothing aligned), q_pos / q_end on the sub-read (after the reverse complement), CIGAR
 * runs with 'I' consuming the sub-read and 'D' the reference, as skbio prints them. */
int lsw_ssw_subreads(lsw_ctx* ctx, int64_t n, const lsw_pair* pairs, const uint8_t* target, int32_t target_len, lsw_result* out,
                     uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used);
int lsw_set_hit_capacity(lsw_ctx* ctx, int64_t cap);   /* 0 = default (total bases / 8 + 4096) */
int lsw_find_all(lsw_ctx* ctx, double identity_threshold, int64_t* n_hits, lsw_counters* counters);
int lsw_fetch_hits(lsw_ctx* ctx, lsw_hit* out, int64_t cap, int64_t* n_out);
int lsw_fetch_hits_compact(lsw_ctx* ctx, lsw_hit16* out, int64_t cap, int64_t* n_out);   /* same hits, 16-byte records */

/* Page-locked host memory for read and hit buffers: copies from / to it run at PCIe speed and asynchronously.  Any host
 * memory works with every entry point; pageable buffers are staged by the driver (several times slower). */
int  lsw_host_alloc(int64_t bytes, void** out);
void lsw_host_free(void* p);

/* Integer-pipe roofline microbenchmark: ops/s of independent chains of one
 * instruction mix on every SM.  mode: see lamprey_b200/csrc/intpeak.cuh.  Returns lane-ops per second. */
int lsw_measure_int_peak(lsw_ctx* ctx, int mode, int iters, double* lane_ops_per_s, float* ms);

int lsw_version(void);

/* ---- Pool: several contexts, one host thread each, behind one handle --------------------------------------------
 * The reference deals the reads of a FASTQ over worker processes and collects their results from a queue
 * (/root/reference/lamprey.py:2182-2207: `read_index % num_processes`, then `q.get()` per worker).  A pool does the same
 * with GPUs: lsw_pool_submit() splits ONE batch into contiguous read ranges balanced by bases, one per device; every range
 * is uploaded, seeded (findAllPrimerAlignments with global read ids) and fetched by its device's worker thread, and the
 * hits of all ranges land in the caller's buffer in global (read, start, query) order -- the host-side gather; there is no
 * device-to-device traffic.  With contexts_per_device = 2, consecutive batches alternate between two contexts of a device,
 * so one batch's PCIe copies overlap another batch's kernels (one lsw_find_all runs on a device at a time).
 * submit returns at once; the read and hit buffers must stay valid until lsw_pool_wait(ticket) returns. */
typedef struct lsw_pool lsw_pool;

#define LSW_HITS_WIDE    0     /* lsw_hit   (32 bytes) */
#define LSW_HITS_COMPACT 1     /* lsw_hit16 (16 bytes) */

typedef struct {
    double upload_s, find_s, fetch_s;   /* host wall clock of the three phases, max over the batch's ranges          */
    double gather_wait_s;               /* longest wait of a range for the hit counts of the ranges before it         */
    double total_s;                     /* submit -> last range done                                                  */
    int64_t reads[8], hits[8];          /* per device (first 8): reads and hits of its range                          */
} lsw_pool_times;

int  lsw_pool_create(int n_devices, const int* devices, int contexts_per_device, lsw_pool** out);
void lsw_pool_destroy(lsw_pool* pool);
const char* lsw_pool_last_error(const lsw_pool* pool);      /* pool may be NULL: error of the last failed lsw_pool_create */
/* scoring + queries + hit test of every context (lsw_set_scoring, lsw_set_queries, lsw_set_overlap, the threshold of
 * lsw_find_all) and the record format of the hit buffers */
int  lsw_pool_configure(lsw_pool* pool, int match, int mismatch, int gap, int gap_ext, int n_queries, const uint8_t* concat,
                        const int32_t* offs, double identity_threshold, int allowed_overlap, int hit_format);
/* the same for the default (scikit-bio StripedSmithWaterman) seeding branch: lsw_set_ssw instead of lsw_set_scoring.  In the 16-byte
 * records of this branch `mismatches` holds the CIGAR columns that are not matches, and the score is not re-derivable: use the wide
 * records when the score is needed. */
int  lsw_pool_configure_ssw(lsw_pool* pool, int match, int mismatch, int gap_open, int gap_extend, int n_queries, const uint8_t* concat,
                            const int32_t* offs, double identity_threshold, int allowed_overlap, int hit_format);
/* pieces > 1: every device's share of a batch is cut further into that many ranges, which alternate between the device's contexts:
 * a single batch then overlaps its own copies with its own kernels (upload of range k+1 and download of range k-1 under the
 * kernels of range k), and a context only needs room for 1/pieces of the share.  Default 1. */
int  lsw_pool_set_pieces(lsw_pool* pool, int pieces);
int  lsw_pool_submit(lsw_pool* pool, int64_t n_reads, const uint8_t* ascii, const int64_t* offs, void* hits_out,
                     int64_t hits_cap, int64_t* ticket);
/* blocks until the batch is done.  *n_hits = hits of the whole batch (> hits_cap with LSW_E_OVERFLOW: nothing was copied
 * for the ranges that did not fit); counters = sums over the ranges, ms_* = device time per device (mean over devices). */
int  lsw_pool_wait(lsw_pool* pool, int64_t ticket, int64_t* n_hits, lsw_counters* counters, lsw_pool_times* times);

/* ---- Host side: FASTQ ingest and the per-read view of a hit list (no CUDA; lsw_hostops.cpp) ---------------------------
 * The reference reads its input with Biopython -- SeqIO.parse(fn, "fastq"), /root/reference/lamprey.py:2145-2155; one delivered
 * file at a time in online mode, lamprey.py:451-507 -- and hands str(record.seq) to every align() call.  lsw_fastq_index scans
 * the bytes of a FASTQ file (4-line records; LF or CRLF; final newline optional; `threads` <= 0: one per core, at most 32) and
 * lsw_fastq_extract writes exactly what lsw_upload_reads / lsw_pool_submit take: the concatenated read text and n + 1 offsets.
 * name_spans (may be NULL) receives [n][2] byte ranges of `data` holding each record's id (the title after '@' up to the first
 * white space: Biopython's record.id).  `data` must stay valid until lsw_fastq_free.  A malformed file (line count not a multiple
 * of 4, no '@' / '+', quality and sequence of different length -- what Biopython raises ValueError for) is LSW_E_ARG with the
 * line or record number in lsw_host_last_error(). */
typedef struct lsw_fastq lsw_fastq;
int  lsw_fastq_index(const uint8_t* data, int64_t size, int threads, lsw_fastq** out, int64_t* n_reads, int64_t* n_bases);
int  lsw_fastq_extract(const lsw_fastq* fq, uint8_t* ascii_out, int64_t* offs_out /* n + 1 */, int64_t* name_spans);
void lsw_fastq_free(lsw_fastq* fq);
/* 16-byte hit records -> 32-byte records (end, score and the survivor flag re-derived as described at lsw_hit16). */
int  lsw_expand_hits(const lsw_hit16* in, int64_t n, int match, int mismatch, int gap, lsw_hit* out, int threads);
/* What findAllPrimerAlignments returns per read next to its hits (lamprey.py:722-731), for a whole sorted hit list of either
 * record format: offsets[r] .. offsets[r + 1] are the hits of read read_base + r (n_reads + 1 entries), covered[r] (may be NULL)
 * is the sum of (end - start) over them -- the numerator of the read's alignment coverage. */
int  lsw_hits_index(const void* hits, int64_t n, int hit_format, int64_t n_reads, int64_t read_base, int64_t* offsets,
                    int64_t* covered, int threads);
const char* lsw_host_last_error(void);                  /* of the calling thread's last failed call of this group */

#ifdef __cplusplus
}
#endif
#endif
