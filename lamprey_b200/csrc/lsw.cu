// lsw.cu -- host side of liblsw.so: the C ABI declared in include/lsw.h.
//
// Drop-in boundary for the seed stage of jackwadden/lamprey:
//   lsw_align_tasks  <->  swalign.LocalAlignment.align(ref, query)        (/root/reference/lamprey.py:563)
//   lsw_find_all     <->  findAllPrimerAlignments / findPrimerAlignments  (/root/reference/lamprey.py:596-645, 699-733)
// The recursion of findPrimerAlignments runs on the device as level-synchronous rounds:
//   round 0:  K1 (score + arg-max of every whole read, block bests stored)  ->  K2 (traceback, hit test, children)
//   round k:  plan / fill the head + tail segments of every child slice  ->  K1 on the segments  ->  merge with the
//             stored block bests (lsw_reuse.cuh)  ->  K2
//   each followed by a scan + scatter of the children into the task list of round k+1.
// No CPU fallback exists: every entry point needs the CUDA device the context was created on.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lsw.h"
#include "lsw_common.cuh"
#include "lsw_score.cuh"
#include "lsw_trace.cuh"
#include "lsw_band.cuh"
#include "lsw_reuse.cuh"
#include "lsw_ssw.cuh"
#include "lsw_affine.cuh"
#include "intpeak.cuh"
#include "lsw_host.hpp"

using namespace lsw;

namespace {

std::string g_create_error;


struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cap = 0; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct TaskBuf {
    DevBuf read, a, b, seq, orig;
    cudaError_t ensure(size_t n, bool with_orig) {
        cudaError_t e;
        if ((e = read.ensure(n * 4)) != cudaSuccess) return e;
        if ((e = a.ensure(n * 4)) != cudaSuccess) return e;
        if ((e = b.ensure(n * 4)) != cudaSuccess) return e;
        if ((e = seq.ensure(n * 2)) != cudaSuccess) return e;
        if (with_orig && (e = orig.ensure(n * 4)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    TaskArrays view() const {
        TaskArrays t;
        t.read = read.as<int32_t>(); t.a = a.as<int32_t>(); t.b = b.as<int32_t>();
        t.seq = seq.as<int16_t>(); t.orig = orig.as<int32_t>(); t.skip = nullptr;
        return t;
    }
    void release() { read.release(); a.release(); b.release(); seq.release(); orig.release(); }
};

struct Timed { cudaEvent_t e0, e1; int kind; };

constexpr int64_t UPLOAD_CHUNK = 64ll << 20;          // bytes of read text per staged host -> device copy (lsw_upload_reads)

}  // namespace

struct lsw_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;               // lsw_upload_reads: host -> staging copies, one chunk ahead of the re-coding
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_coded[2] = {nullptr, nullptr};
    std::string err;
    int num_sms = 0;

    int match = 2, mismatch = -1, g = 1;
    // the default (skbio / SSW) seeding path (lsw_ssw.cuh): on after lsw_set_ssw(), off after lsw_set_scoring()
    bool ssw = false;
    int ssw_match = 2, ssw_mismatch = -3, ssw_open = 5, ssw_extend = 2;
    int nseq = 0;
    std::vector<int32_t> qlen;
    std::vector<uint8_t> qcodes;
    std::vector<int32_t> class_of;
    std::vector<int32_t> class_seqs[NCLASS];
    DevBuf d_qcodes, d_qlen, d_class_seqs[NCLASS], d_cand_ok, d_wx;

    int64_t n_reads = 0, total_bases = 0, codes_bytes = 0;
    std::vector<int32_t> h_rlen;
    DevBuf d_ascii, d_offs_in, d_codes, d_roff, d_rlen, d_order, d_order_tmp, d_len_tmp[2];

    TaskBuf tb[2], child;
    DevBuf d_fb2[NCLASS], d_fb2cnt, d_head2;
    DevBuf d_fallback[NCLASS], d_nfallback, d_cand_key[NCLASS], d_cand_sorted[NCLASS], d_cand_key_sorted[NCLASS], d_cand_rec[NCLASS];
    DevBuf d_res, d_seg[2], d_head, d_cand[NCLASS], d_ncand, d_counters, d_child_flag, d_pos,
           d_cubtemp, d_scratch, d_hits, d_hits_sorted, d_hits16, d_nhits, d_keys[2], d_vals[2], d_err,
           d_results, d_cigar, d_ncigar;
    int64_t hit_cap_user = 0;
    int32_t read_base = 0;
    // reuse of the round-0 DP by child tasks (lsw_reuse.cuh)
    bool reuse = true;
    std::vector<int64_t> h_bb_off, h_bb1_off;
    int64_t bb_total = 0, bb1_total = 0;
    bool use_bb1 = false;        // second-level block bests: built when the reads are long enough for slices to span whole groups
    DevBuf d_blockbest, d_bb_off, d_bb1, d_bb1_off, d_nseg, d_segpos, d_res_s, d_seg_s, d_class_of, d_kt;
    int32_t allowed_overlap = -1;
    int64_t n_hits = 0;
    bool hits_pruned = false;
    // environment switches, read once in lsw_create
    int env_k2_blocks = 0, env_k2_debug = 0;
    bool env_debug = false, env_no_reuse = false, env_ssw_scalar = false, env_full_head = false;        // the sorted hits carry the survivor flag of lsw_set_overlap (lsw_chain_subreads needs it)
    std::vector<Timed> timed;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

namespace {

// every device buffer of a context (lsw_destroy releases them; lsw_counters.device_bytes sums their capacities)
std::vector<DevBuf*> all_bufs(lsw_ctx* ctx) {
    std::vector<DevBuf*> v = {&ctx->d_qcodes, &ctx->d_qlen, &ctx->d_cand_ok, &ctx->d_wx, &ctx->d_ascii, &ctx->d_offs_in, &ctx->d_codes,
                              &ctx->d_roff, &ctx->d_rlen, &ctx->d_order, &ctx->d_order_tmp, &ctx->d_len_tmp[0], &ctx->d_len_tmp[1],
                              &ctx->d_res, &ctx->d_seg[0], &ctx->d_seg[1], &ctx->d_head, &ctx->d_ncand, &ctx->d_counters,
                              &ctx->d_child_flag, &ctx->d_pos, &ctx->d_cubtemp, &ctx->d_scratch, &ctx->d_hits, &ctx->d_hits_sorted,
                              &ctx->d_hits16, &ctx->d_nhits, &ctx->d_keys[0], &ctx->d_keys[1], &ctx->d_vals[0], &ctx->d_vals[1],
                              &ctx->d_err, &ctx->d_results, &ctx->d_cigar, &ctx->d_ncigar, &ctx->d_nfallback, &ctx->d_fb2cnt,
                              &ctx->d_head2, &ctx->d_blockbest, &ctx->d_bb_off, &ctx->d_bb1, &ctx->d_bb1_off, &ctx->d_nseg,
                              &ctx->d_segpos, &ctx->d_res_s, &ctx->d_seg_s, &ctx->d_class_of, &ctx->d_kt};
    for (int c = 0; c < NCLASS; c++)
        for (DevBuf* b : {&ctx->d_class_seqs[c], &ctx->d_cand[c], &ctx->d_fallback[c], &ctx->d_cand_key[c], &ctx->d_cand_sorted[c],
                          &ctx->d_cand_key_sorted[c], &ctx->d_cand_rec[c], &ctx->d_fb2[c]}) v.push_back(b);
    for (TaskBuf* t : {&ctx->tb[0], &ctx->tb[1], &ctx->child})
        for (DevBuf* b : {&t->read, &t->a, &t->b, &t->seq, &t->orig}) v.push_back(b);
    return v;
}

int64_t device_bytes(lsw_ctx* ctx) {
    int64_t n = 0;
    for (DevBuf* b : all_bufs(ctx)) n += (int64_t)b->cap;
    return n;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char buf_[512];                                                                   \
            snprintf(buf_, sizeof buf_, "%s:%d: %s: %s", __FILE__, __LINE__, #call,            \
                     cudaGetErrorString(e_));                                                 \
            ctx->err = buf_;                                                                  \
            return LSW_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

int fail(lsw_ctx* ctx, int code, const char* msg) { ctx->err = msg; return code; }

// the environment switches, once per call of an entry point (not per launch: tests flip them between calls on one context)
void read_env(lsw_ctx* ctx) {
    const char* e;
    ctx->env_k2_blocks = (e = getenv("LSW_K2_BLOCKS_PER_SM")) ? atoi(e) : 0;
    ctx->env_k2_debug = (e = getenv("LSW_K2_DEBUG")) ? atoi(e) : 0;
    ctx->env_debug = getenv("LSW_DEBUG") != nullptr;
    ctx->env_no_reuse = getenv("LSW_NO_REUSE") != nullptr;
    ctx->env_ssw_scalar = getenv("LSW_SSW_SCALAR") != nullptr;
    ctx->env_full_head = getenv("LSW_REUSE_FULL_HEAD") != nullptr;
}

// ---- small kernels ----------------------------------------------------------------------

// ASCII -> read codes (A,C,G,T -> 0..3 after upper-casing, anything else -> 4), one warp per read,
// each read written at a 16-byte aligned offset.
// `ascii` is a staged chunk of the batch: the bases of reads [r0, r1), whose first byte is the caller's offset `base`.
__global__ void k_encode_reads(const uint8_t* __restrict__ ascii, const int64_t* __restrict__ offs_in, int64_t base,
                               const int64_t* __restrict__ roff, uint8_t* __restrict__ codes, int64_t r0, int64_t r1) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = r0 + warp; i < r1; i += nwarps) {
        const int64_t s = offs_in[i] - base, len = offs_in[i + 1] - offs_in[i], d = roff[i];
        const int64_t padded = (len + 15) & ~15ll;
        for (int64_t j = lane; j < padded; j += 32) {
            const uint8_t code = j < len ? encode_base(ascii[s + j]) : (uint8_t)CODE_OTHER;
            codes[d + j] = code;
        }
    }
}

// Round 0 of findAllPrimerAlignments: every sequence against every whole read.
// Within a sequence the reads are queued longest first (order[]: indices sorted by descending length), so
// the few very long reads of an ONT length distribution start early instead of forming the tail of the launch.
__global__ void k_init_tasks(TaskArrays t, const int32_t* __restrict__ rlen, const int32_t* __restrict__ order,
                             int64_t n_reads, int nseq) {
    const int64_t n = n_reads * nseq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / n_reads);
        const int32_t r = order[i - (int64_t)s * n_reads];
        t.read[i] = r; t.a[i] = 0; t.b[i] = rlen[r]; t.seq[i] = (int16_t)s;
    }
}

struct PadLen {
    __host__ __device__ long long operator()(const int32_t& len) const { return ((long long)len + 15) & ~15ll; }
};

__global__ void k_scatter_offsets(const int32_t* __restrict__ order, const long long* __restrict__ pos,
                                  int64_t* __restrict__ roff, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        roff[order[i]] = 256 + pos[i];
}

// start descriptors of a plain task list (every task is a whole slice, no warm-up)
__global__ void k_make_ktasks(TaskArrays t, const int64_t* __restrict__ roff, KTask* __restrict__ kt, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        KTask k; k.p0 = roff[t.read[i]] + t.a[i]; k.len = t.b[i] - t.a[i]; k.skip = 0;
        kt[i] = k;
    }
}

__global__ void k_iota(int32_t* __restrict__ v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (int32_t)i;
}

struct FlagToInt {
    __host__ __device__ int32_t operator()(const uint8_t& f) const { return (int32_t)f; }
};

// children of round k (slots 2t, 2t+1) -> dense task list of round k+1; order is preserved, so the
// tasks of one sequence stay contiguous and seg_next[s] = pos[2 * seg[s]].
__global__ void k_scatter_children(TaskArrays child, const uint8_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                   int64_t n2, TaskArrays next, const int64_t* __restrict__ seg, int64_t* __restrict__ seg_next,
                                   int nseq) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid <= nseq) {
        const int64_t j = 2 * seg[gid];
        seg_next[gid] = pos[j < n2 ? j : n2];
    }
    for (int64_t j = gid; j < n2; j += (int64_t)gridDim.x * blockDim.x) {
        if (flag[j]) {
            const int32_t d = pos[j];
            next.read[d] = child.read[j]; next.a[d] = child.a[j]; next.b[d] = child.b[j]; next.seq[d] = child.seq[j];
        }
    }
}

// sort key of a hit: (read, start, query) -- the order lamprey.py:722's stable sort yields
// (the fields are packed as tightly as the batch allows: every 8 key bits saved is one radix pass over all hits)
__global__ void k_hit_keys(const lsw_hit* __restrict__ hits, int64_t n, unsigned long long* __restrict__ keys,
                           uint32_t* __restrict__ vals, int qbits, int sbits) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const lsw_hit h = hits[i];
        keys[i] = ((unsigned long long)(uint32_t)h.read << (qbits + sbits)) | ((unsigned long long)(uint32_t)h.start << qbits) |
                  (unsigned long long)(uint16_t)h.query;
        vals[i] = (uint32_t)i;
    }
}

__global__ void k_gather_hits(const lsw_hit* __restrict__ hits, const uint32_t* __restrict__ vals, int64_t n,
                              lsw_hit* __restrict__ out, int32_t read_base) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int4* src = reinterpret_cast<const int4*>(hits + vals[i]);
        int4* dst = reinterpret_cast<int4*>(out + i);
        int4 a = src[0];
        a.x += read_base;                   // lsw_hit.read: index in the caller's whole batch (lsw_set_read_base)
        dst[0] = a; dst[1] = src[1];
    }
}

// sorted 32-byte hits -> 16-byte records (lsw_hit16): one 16-byte store per hit
__global__ void k_pack_hits16(const lsw_hit* __restrict__ hits, int64_t n, uint4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int4* src = reinterpret_cast<const int4*>(hits + i);
        const int4 a = src[0], b = src[1];          // a: read, start, end, query | depth << 16;  b: m | x << 16, score | q_pos << 16, q_end | reserved << 16
        const uint32_t query = (uint32_t)a.w & 0xFFFFu, depth = (uint32_t)a.w >> 16;
        const uint32_t m = (uint32_t)b.x & 0xFFFFu, x = (uint32_t)b.x >> 16;
        const uint32_t q_pos = (uint32_t)b.y >> 16, q_end = (uint32_t)b.z & 0xFFFFu, keep = ((uint32_t)b.z >> 16) & 1u;
        uint4 o;
        o.x = (uint32_t)a.x; o.y = (uint32_t)a.y;
        o.z = ((uint32_t)(a.z - a.y) & 0xFFu) | ((query & 0xFFu) << 8) | ((depth > 0x7FFFu ? 0xFFFFu : depth) << 16);
        o.w = (m & 0x7Fu) | (keep << 7) | ((x & 0xFFu) << 8) | ((q_pos & 0xFFu) << 16) | ((q_end & 0xFFu) << 24);
        out[i] = o;
    }
}

// one thread per read: find the read's segment in the sorted hits and prune overlaps (lsw_set_overlap)
__device__ __forceinline__ void hit_range_of_read(const lsw_hit* __restrict__ hits, int64_t n_hits, int32_t rid, long long& lo, long long& e) {
    lo = 0;
    long long hi = n_hits;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (hits[mid].read < rid) lo = mid + 1; else hi = mid; }
    e = lo;
    long long eh = n_hits;
    while (e < eh) { const long long mid = (e + eh) >> 1; if (hits[mid].read <= rid) e = mid + 1; else eh = mid; }
}

__global__ void k_prune_overlaps(lsw_hit* __restrict__ hits, int64_t n_hits, int64_t n_reads, int32_t read_base, int allowed,
                                 const int32_t* __restrict__ qlen /* skbio path: identity = matches / len(primer); else null */) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += (int64_t)gridDim.x * blockDim.x) {
        long long lo, e;
        hit_range_of_read(hits, n_hits, (int32_t)i + read_base, lo, e);
        prune_overlaps_read(hits, lo, e, allowed, qlen);
    }
}

// stage 1 of processLamplicon (lsw_chain.cuh), one thread per read: pass 1 counts the targets of every read, pass 2 (after a
// scan) writes the sub-reads
__global__ void k_chain_subreads(const lsw_hit* __restrict__ hits, int64_t n_hits, int64_t n_reads, int32_t read_base,
                                 const int32_t* __restrict__ rlen, ChainOrders C, int32_t* __restrict__ count,
                                 const int32_t* __restrict__ pos, lsw_subread* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_reads; i += (int64_t)gridDim.x * blockDim.x) {
        long long lo, e;
        hit_range_of_read(hits, n_hits, (int32_t)i + read_base, lo, e);
        if (!out) count[i] = chain_subreads_read(hits, lo, e, rlen[i], C, nullptr);
        else chain_subreads_read(hits, lo, e, rlen[i], C, out + pos[i]);
    }
}

// ---- reuse of the round-0 DP (lsw_reuse.cuh): plan, fill, merge ----------------------------------------
__global__ void k_plan_segments(ReuseParams R) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= R.n; t += (int64_t)gridDim.x * blockDim.x)
        R.nseg[t] = t < R.n ? plan_of(R, t).nseg : 0;
}

__global__ void k_fill_segments(ReuseParams R, const int64_t* __restrict__ seg_t, int64_t* __restrict__ seg_s, int nseq) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid <= nseq) seg_s[gid] = R.segpos[seg_t[gid]];
    for (int64_t t = gid; t < R.n; t += (int64_t)gridDim.x * blockDim.x) fill_segments(R, t);
}

// second-level block bests (lsw_reuse.cuh): one thread per (sequence, read, group of BB_GROUP blocks)
__global__ void k_build_bb1(const uint32_t* __restrict__ bb, const int64_t* __restrict__ bb_off, int64_t bb_total,
                            uint32_t* __restrict__ bb1, const int64_t* __restrict__ bb1_off, int64_t bb1_total,
                            int64_t n_reads, int nseq, int g) {
    const int64_t n = bb1_total * nseq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i / bb1_total);
        const int64_t G1 = i - (int64_t)q * bb1_total;            // group index over all reads
        int64_t lo = 0, hi = n_reads;                              // read that owns it: last rd with bb1_off[rd] <= G1
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (bb1_off[mid] <= G1) lo = mid; else hi = mid; }
        const int64_t rd = lo;
        bb1[i] = reduce_block_group(bb + (size_t)q * bb_total + bb_off[rd], (int)(bb_off[rd + 1] - bb_off[rd]), (int)(G1 - bb1_off[rd]), g);
    }
}

__global__ void k_merge_segments(ReuseParams R) {
    unsigned long long tasks = 0, cells = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < R.n; t += (int64_t)gridDim.x * blockDim.x) {
        int score, row, col;
        const bool cand = combine_task(R, t, score, row, col);
        const int q = R.t.seq[t];
        tasks += 1;
        cells += (unsigned long long)(R.t.b[t] - R.t.a[t]) * (unsigned long long)R.qlen[q];
        if (cand) {
            const int c = R.class_of[q];
            const unsigned long long i = append_slot_grouped(R.ncand + c);
            CandRec cr;
            cr.rd = R.t.read[t]; cr.a = R.t.a[t]; cr.len = R.t.b[t] - cr.a; cr.p0 = R.roff[cr.rd] + cr.a;
            cr.res_x = score | (row << 16); cr.col = col; cr.t = (int32_t)t;
            R.cand[c][i] = (uint32_t)i;
            R.cand_key[c][i] = (uint16_t)(((uint32_t)q << 8) | (uint32_t)cand_window(score, row, col, R.match, R.g));
            R.cand_rec[c][i] = cr;
        }
    }
    // block-level reduction of the work counters
    __shared__ unsigned long long s_tasks, s_cells;
    if (threadIdx.x == 0) { s_tasks = 0; s_cells = 0; }
    __syncthreads();
    atomicAdd(&s_tasks, tasks); atomicAdd(&s_cells, cells);
    __syncthreads();
    if (threadIdx.x == 0 && s_tasks) { atomicAdd(&R.counters[0], s_tasks); atomicAdd(&R.counters[1], s_cells); }
}

// ---- helpers -----------------------------------------------------------------------------

// lsw_measure_int_peak mode 9: K1's inner loop and nothing else -- ScoreLane<24>::column2 (per row and column pair: 2 VIADDMNMX,
// 3 VIMNMX3, 4 adds, one LDS.64 of the profile) on a profile in shared memory, with K1's block size, register budget and
// occupancy.  What K1 would reach if task switches, chunk loads, masks and the arg-max folds were free.
template <int KEYV>
__global__ void __launch_bounds__(SCORE_WARPS * 32, ScoreOcc<24>::value)
k_intpeak_k1row(uint32_t* out, const uint32_t* in, int iters) {
    using Lane = ScoreLane<24>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    U2* prof = reinterpret_cast<U2*>(smem_raw);
    uint32_t* p32 = reinterpret_cast<uint32_t*>(prof);
    for (int e = threadIdx.x; e < NIDX * Lane::PROF_STRIDE * 2; e += blockDim.x) p32[e] = ((e * 7 + in[0]) & 1) ? 0x03000300u : 0x00000000u;
    __syncthreads();
    Lane L;
    L.init();
    const uint32_t G2 = 0x01000100u;
    uint32_t tvec = 0, ia = (threadIdx.x + in[1]) & 15u, ib = (threadIdx.x * 3 + in[2]) & 15u;
    for (int it = 0; it < iters; it++) {
        L.template column2<KEYV>(prof + ia * Lane::PROF_STRIDE, prof + ib * Lane::PROF_STRIDE, tvec, tvec + 0x00010001u, G2, G2, in[5]);
        tvec = (tvec + 0x00020002u) & 0x003F003Fu;
        ia = (ia + 5u) & 15u; ib = (ib + 3u) & 15u;
    }
    uint32_t sum = 0;
#pragma unroll
    for (int r = 0; r < 24; r++) sum += L.M[r];
#pragma unroll
    for (int r = 0; r < 12; r++) sum += L.K[r];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

int grid_for(int64_t n, int threads, int cap_blocks) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap_blocks) b = cap_blocks;
    return (int)b;
}

cudaEvent_t next_event(lsw_ctx* ctx) {
    if (ctx->ev_used == ctx->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_pool[ctx->ev_used++];
}

struct Span {      // times a region of the stream into one of the lsw_counters buckets
    lsw_ctx* ctx; Timed t;
    Span(lsw_ctx* c, int kind) : ctx(c) {
        t.kind = kind; t.e0 = next_event(c); t.e1 = next_event(c);
        if (t.e0) cudaEventRecord(t.e0, c->stream);
    }
    ~Span() { if (t.e1) { cudaEventRecord(t.e1, ctx->stream); ctx->timed.push_back(t); } }
};

// Launch configuration of one kernel instantiation (dynamic shared memory attribute + resident blocks per SM), looked up once per
// process and device: a step launches ~560 kernels, most of them in the late, tiny recursion rounds where the GPU waits for the
// host -- two runtime calls per launch were a measurable part of a step there.
struct LaunchCfg {
    std::atomic<int> occ[16];
    LaunchCfg() { for (auto& o : occ) o.store(-1); }
    template <class K>
    cudaError_t get(int device, K kernel, int threads, size_t smem, int& out) {
        const int d = device & 15;
        int v = occ[d].load(std::memory_order_acquire);
        if (v < 0) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem);
            if (e != cudaSuccess) return e;
            occ[d].store(v, std::memory_order_release);
        }
        out = v;
        return cudaSuccess;
    }
};

template <int NR>
int launch_score(lsw_ctx* ctx, const ScoreParams& P, int64_t n_class_tasks) {
    const size_t smem = score_smem_bytes<NR>();
    static LaunchCfg cfg;
    int occ = 0;
    CK(cfg.get(ctx->device, k_score<NR>, SCORE_WARPS * 32, smem, occ));
    if (occ < 1) return fail(ctx, LSW_E_INTERNAL, "score kernel does not fit on an SM");
    const int64_t lanes_needed = (n_class_tasks + 1) / 2;
    const int blocks = grid_for(lanes_needed, SCORE_WARPS * 32, ctx->num_sms * occ);
    k_score<NR><<<blocks, SCORE_WARPS * 32, smem, ctx->stream>>>(P);
    CK(cudaGetLastError());
    return LSW_OK;
}

int launch_score_class(lsw_ctx* ctx, int c, const ScoreParams& P, int64_t n) {
    switch (CLASS_NR[c]) {
        case 16: return launch_score<16>(ctx, P, n);
        case 20: return launch_score<20>(ctx, P, n);
        case 24: return launch_score<24>(ctx, P, n);
        case 28: return launch_score<28>(ctx, P, n);
        case 32: return launch_score<32>(ctx, P, n);
        case 40: return launch_score<40>(ctx, P, n);
        case 48: return launch_score<48>(ctx, P, n);
        case 56: return launch_score<56>(ctx, P, n);
        default: return launch_score<64>(ctx, P, n);
    }
}

#ifndef LSW_BAND_WORDS
#define LSW_BAND_WORDS 2       // 16-row band (lsw_band.cuh): ~1 % of the walks leave it and take the full-window kernel; 8 rows: ~19 %
#endif

template <int NR, int NR2>
int launch_trace(lsw_ctx* ctx, TraceParams P, int64_t n_class_tasks, int qmax) {
    // tier 1, every candidate of the class: one lane per PAIR of candidates, window bands in shared memory (lsw_band.cuh)
    constexpr int BW = LSW_BAND_WORDS;
    const size_t smem = BandSmem<NR, BW>::total;
    // find_all never needs the op list of a path, align_tasks (CIGAR output) does
    auto kernel = P.mode == 1 ? k_trace_band<NR, BW, true> : k_trace_band<NR, BW, false>;
    static LaunchCfg cfg_band[2], cfg_fast[2], cfg_full;
    int occ = 0;
    CK(cfg_band[P.mode == 1].get(ctx->device, kernel, TRACE_THREADS, smem, occ));
    if (occ < 1) return fail(ctx, LSW_E_INTERNAL, "trace kernel does not fit on an SM");
    if (ctx->env_k2_blocks > 0) occ = std::max(1, std::min(occ, ctx->env_k2_blocks));
    const int blocks = grid_for((n_class_tasks + 1) / 2, TRACE_THREADS, ctx->num_sms * occ);
    // tier 2, the ~1 % whose walk left the band: the whole window in a global scratch (k_trace_fast); a small grid, the list is short
    const size_t smem1 = trace_smem_bytes<NR>();
    auto kernel1 = P.mode == 1 ? k_trace_fast<NR, true> : k_trace_fast<NR, false>;
    int occ1 = 0;
    CK(cfg_fast[P.mode == 1].get(ctx->device, kernel1, TRACE_THREADS, smem1, occ1));
    const int blocks1 = grid_for((n_class_tasks + 1) / 2, TRACE_THREADS, ctx->num_sms * 2);
    const int64_t threads1 = (int64_t)blocks1 * TRACE_THREADS;
    const int64_t Wf = trace_window_fast_max(qmax, ctx->match, ctx->g);
    constexpr int UNITS = TraceLane<NR>::UPAIRS;         // 16 bytes per (window column, unit pair, lane): both tasks of the lane
    // tier 3, what even that could not prove: one thread per task, 2-bit ops of a window that is exact a priori
    int occ2 = 0;
    CK(cfg_full.get(ctx->device, k_trace_full<NR2>, TRACE_THREADS, 0, occ2));
    if (occ2 < 1) return fail(ctx, LSW_E_INTERNAL, "trace kernel does not fit on an SM");
    const int blocks2 = grid_for(n_class_tasks, TRACE_THREADS, ctx->num_sms * std::min(occ2, 2));
    const int64_t threads2 = (int64_t)blocks2 * TRACE_THREADS;
    const int64_t W = trace_window_full(qmax, ctx->match, ctx->g);
    if ((size_t)threads2 * (size_t)W * (NR2 / 16) >= (1ull << 31) || (size_t)threads1 * (size_t)(Wf + 4) * UNITS * 4 >= (1ull << 31))
        return fail(ctx, LSW_E_INTERNAL, "traceback scratch exceeds the 32-bit index range");
    const size_t need = std::max((size_t)threads1 * (size_t)(Wf + 4) * UNITS * 16, (size_t)threads2 * (size_t)W * (NR2 / 16) * 4);
    CK(ctx->d_scratch.ensure(need));
    P.scratch = ctx->d_scratch.as<uint32_t>();
    kernel<<<blocks, TRACE_THREADS, smem, ctx->stream>>>(P);
    CK(cudaGetLastError());
    TraceParams P1 = P;
    P1.head = ctx->d_head2.as<unsigned long long>();
    P1.fb_in = P.fb2; P1.fb_in_cnt = P.fb2_cnt;
    P1.scratch_stride = TRACE_THREADS;
    P1.scratch_block_words = (int64_t)(Wf + 4) * UNITS * TRACE_THREADS;
    kernel1<<<blocks1, TRACE_THREADS, smem1, ctx->stream>>>(P1);
    CK(cudaGetLastError());
    P.scratch_stride = threads2;
    k_trace_full<NR2><<<blocks2, TRACE_THREADS, 0, ctx->stream>>>(P);
    CK(cudaGetLastError());
    return LSW_OK;
}

int launch_trace_class(lsw_ctx* ctx, int c, const TraceParams& P, int64_t n, int qmax) {
    switch (CLASS_NR[c]) {
        case 16: return launch_trace<16, 16>(ctx, P, n, qmax);
        case 20: return launch_trace<20, 32>(ctx, P, n, qmax);
        case 24: return launch_trace<24, 32>(ctx, P, n, qmax);
        case 28: return launch_trace<28, 32>(ctx, P, n, qmax);
        case 32: return launch_trace<32, 32>(ctx, P, n, qmax);
        case 40: return launch_trace<40, 48>(ctx, P, n, qmax);
        case 48: return launch_trace<48, 48>(ctx, P, n, qmax);
        case 56: return launch_trace<56, 64>(ctx, P, n, qmax);
        default: return launch_trace<64, 64>(ctx, P, n, qmax);
    }
}

void fill_score_params(lsw_ctx* ctx, ScoreParams& P, const TaskBuf& cur, int segi) {
    memset(&P, 0, sizeof P);
    P.codes = ctx->d_codes.as<uint8_t>();
    P.roff = ctx->d_roff.as<int64_t>();
    P.t = cur.view();
    P.seg = ctx->d_seg[segi].as<int64_t>();
    P.head = ctx->d_head.as<unsigned long long>();
    P.res = ctx->d_res.as<int2>();
    P.qcodes = ctx->d_qcodes.as<uint8_t>();
    P.qlen = ctx->d_qlen.as<int32_t>();
    P.cand_ok = ctx->d_cand_ok.as<uint8_t>();
    P.counters = ctx->d_counters.as<unsigned long long>();
    P.s_match = SC * (ctx->match + ctx->g);
    P.s_mis = SC * (ctx->mismatch + ctx->g);
    P.g = ctx->g;
    P.match = ctx->match;
}

void fill_trace_params(lsw_ctx* ctx, TraceParams& T, const TaskBuf& cur) {
    memset(&T, 0, sizeof T);
    T.codes = ctx->d_codes.as<uint8_t>();
    T.roff = ctx->d_roff.as<int64_t>();
    T.t = cur.view();
    T.res = ctx->d_res.as<int2>();
    T.qcodes = ctx->d_qcodes.as<uint8_t>();
    T.qlen = ctx->d_qlen.as<int32_t>();
    T.cand_ok = ctx->d_cand_ok.as<uint8_t>();
    T.match = ctx->match; T.mismatch = ctx->mismatch; T.g = ctx->g;
    T.errflag = ctx->d_err.as<int32_t>();
    T.all_fallback = nibble_walk_ok(ctx->match, ctx->mismatch, ctx->g) ? 0 : 1;
    T.debug = ctx->env_k2_debug;
}

// Runs K1 then K2 for every class over the task list `cur` (segments h_seg on the host).
int run_round(lsw_ctx* ctx, const TaskBuf& cur, int segi, const std::vector<int64_t>& h_seg, int64_t n,
              TraceParams Tbase, lsw_counters* cnt, int round_kind = 0 /* 0 plain, 1 store block bests, 2 reuse them */) {
    CK(ctx->d_res.ensure((size_t)n * sizeof(int2)));
    CK(ctx->d_head.ensure((size_t)ctx->nseq * 8));
    CK(ctx->d_ncand.ensure(NCLASS * 8));
    CK(cudaMemsetAsync(ctx->d_head.p, 0, (size_t)ctx->nseq * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ncand.p, 0, NCLASS * 8, ctx->stream));
    CK(ctx->d_nfallback.ensure(NCLASS * 8));
    CK(cudaMemsetAsync(ctx->d_nfallback.p, 0, NCLASS * 8, ctx->stream));
    CK(ctx->d_fb2cnt.ensure((size_t)ctx->nseq * 8)); CK(ctx->d_head2.ensure((size_t)ctx->nseq * 8));
    CK(cudaMemsetAsync(ctx->d_fb2cnt.p, 0, (size_t)ctx->nseq * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_head2.p, 0, (size_t)ctx->nseq * 8, ctx->stream));
    int64_t ncls[NCLASS];
    int qmax[NCLASS];
    for (int c = 0; c < NCLASS; c++) {
        ncls[c] = 0; qmax[c] = 0;
        for (int s : ctx->class_seqs[c]) { ncls[c] += h_seg[s + 1] - h_seg[s]; qmax[c] = std::max(qmax[c], ctx->qlen[s]); }
        if (ncls[c]) {
            CK(ctx->d_cand[c].ensure((size_t)ncls[c] * 4)); CK(ctx->d_fallback[c].ensure((size_t)ncls[c] * 4)); CK(ctx->d_fb2[c].ensure((size_t)ncls[c] * 4));
            CK(ctx->d_cand_sorted[c].ensure((size_t)ncls[c] * 4));
            CK(ctx->d_cand_rec[c].ensure((size_t)ncls[c] * sizeof(CandRec)));
            CK(ctx->d_cand_key[c].ensure((size_t)ncls[c] * 2)); CK(ctx->d_cand_key_sorted[c].ensure((size_t)ncls[c] * 2));
            CK(cudaMemsetAsync(ctx->d_cand_key[c].p, 0xFF, (size_t)ncls[c] * 2, ctx->stream));
        }
    }
    ScoreParams P;
    fill_score_params(ctx, P, cur, segi);
    P.push_cands = 1; P.count_ref = 1; P.block_steps = BLOCK_STEPS;
    if (round_kind == 1) {
        P.blockbest = ctx->d_blockbest.as<uint32_t>(); P.bb_off = ctx->d_bb_off.as<int64_t>(); P.bb_total = ctx->bb_total;
    }
    ReuseParams R;
    memset(&R, 0, sizeof R);
    if (round_kind != 2) {
        Span sp(ctx, 2);
        CK(ctx->d_kt.ensure((size_t)n * sizeof(KTask)));
        k_make_ktasks<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(cur.view(), ctx->d_roff.as<int64_t>(), ctx->d_kt.as<KTask>(), n);
        CK(cudaGetLastError());
        P.kt = ctx->d_kt.as<KTask>();
        if (cnt) cnt->other_launches++;
    }
    if (round_kind == 2) {
        // child tasks: K1 runs on the segment list (own head + warm-started tail), the whole-read blocks in between are looked up
        Span sp(ctx, 2);
        R.t = cur.view(); R.n = n; R.qlen = ctx->d_qlen.as<int32_t>(); R.match = ctx->match; R.g = ctx->g; R.enable = 1;
        R.wx = ctx->d_wx.as<int32_t>();
        R.blockbest = ctx->d_blockbest.as<uint32_t>(); R.bb_off = ctx->d_bb_off.as<int64_t>(); R.bb_total = ctx->bb_total;
        if (ctx->use_bb1) { R.bb1 = ctx->d_bb1.as<uint32_t>(); R.bb1_off = ctx->d_bb1_off.as<int64_t>(); R.bb1_total = ctx->bb1_total; }
        CK(ctx->d_nseg.ensure((size_t)(n + 1) * 4)); CK(ctx->d_segpos.ensure((size_t)(n + 1) * 4));
        CK(ctx->d_res_s.ensure((size_t)2 * n * sizeof(int2))); CK(ctx->d_seg_s.ensure((size_t)(ctx->nseq + 1) * 8));
        R.nseg = ctx->d_nseg.as<int32_t>(); R.segpos = ctx->d_segpos.as<int32_t>();
        CK(ctx->d_kt.ensure((size_t)2 * n * sizeof(KTask)));
        R.s_kt = ctx->d_kt.as<KTask>(); R.roff = ctx->d_roff.as<int64_t>();
        k_plan_segments<<<grid_for(n + 1, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(R);
        CK(cudaGetLastError());
        size_t tmp = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_nseg.as<int32_t>(), ctx->d_segpos.as<int32_t>(), (int)(n + 1), ctx->stream));
        CK(ctx->d_cubtemp.ensure(tmp));
        CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtemp.p, tmp, ctx->d_nseg.as<int32_t>(), ctx->d_segpos.as<int32_t>(), (int)(n + 1), ctx->stream));
        k_fill_segments<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            R, ctx->d_seg[segi].as<int64_t>(), ctx->d_seg_s.as<int64_t>(), ctx->nseq);
        CK(cudaGetLastError());
        if (cnt) cnt->other_launches += 4;
        P.t = TaskArrays(); P.kt = ctx->d_kt.as<KTask>(); P.seg = ctx->d_seg_s.as<int64_t>(); P.res = ctx->d_res_s.as<int2>();
        P.push_cands = 0; P.count_ref = 0; P.block_steps = SEG_BLOCK;
    }
    {
        Span sp(ctx, (Tbase.mode == 0 && Tbase.round == 0) ? 4 : 0);      // 4: K1 on whole reads (round 0 of find_all)
        for (int c = 0; c < NCLASS; c++) {
            if (!ncls[c]) continue;
            P.class_seqs = ctx->d_class_seqs[c].as<int32_t>();
            P.n_class_seqs = (int)ctx->class_seqs[c].size();
            P.cand = ctx->d_cand[c].as<uint32_t>();
            P.cand_key = ctx->d_cand_key[c].as<uint16_t>();
            P.cand_rec = ctx->d_cand_rec[c].as<CandRec>();
            P.ncand = ctx->d_ncand.as<unsigned long long>() + c;
            int rc = launch_score_class(ctx, c, P, round_kind == 2 ? 2 * ncls[c] : ncls[c]);
            if (rc) return rc;
            if (cnt) cnt->score_launches++;
        }
    }
    if (round_kind == 2) {
        Span sp(ctx, 2);
        R.res_s = ctx->d_res_s.as<int2>(); R.res_t = ctx->d_res.as<int2>();
        R.blockbest = ctx->d_blockbest.as<uint32_t>(); R.bb_off = ctx->d_bb_off.as<int64_t>(); R.bb_total = ctx->bb_total;
        R.cand_ok = ctx->d_cand_ok.as<uint8_t>(); R.class_of = ctx->d_class_of.as<int32_t>();
        for (int c = 0; c < NCLASS; c++) { R.cand[c] = ctx->d_cand[c].as<uint32_t>(); R.cand_key[c] = ctx->d_cand_key[c].as<uint16_t>(); R.cand_rec[c] = ctx->d_cand_rec[c].as<CandRec>(); }
        R.ncand = ctx->d_ncand.as<unsigned long long>();
        R.counters = ctx->d_counters.as<unsigned long long>();
        k_merge_segments<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(R);
        CK(cudaGetLastError());
        if (cnt) cnt->other_launches++;
    }
    {
        // order each class's candidates by (sequence, traceback window length): K2 warps serve one sequence at a
        // time and their lanes then fill windows of similar length
        Span sp(ctx, 2);
        for (int c = 0; c < NCLASS; c++) {
            if (!ncls[c]) continue;
            size_t tmp = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->d_cand_key[c].as<uint16_t>(), ctx->d_cand_key_sorted[c].as<uint16_t>(),
                                               ctx->d_cand[c].as<uint32_t>(), ctx->d_cand_sorted[c].as<uint32_t>(), (int)ncls[c], 0, 16, ctx->stream));
            CK(ctx->d_cubtemp.ensure(tmp));
            CK(cub::DeviceRadixSort::SortPairs(ctx->d_cubtemp.p, tmp, ctx->d_cand_key[c].as<uint16_t>(), ctx->d_cand_key_sorted[c].as<uint16_t>(),
                                               ctx->d_cand[c].as<uint32_t>(), ctx->d_cand_sorted[c].as<uint32_t>(), (int)ncls[c], 0, 16, ctx->stream));
            if (cnt) cnt->other_launches += 2;
        }
        CK(cudaMemsetAsync(ctx->d_head.p, 0, (size_t)ctx->nseq * 8, ctx->stream));     // K2 reuses the per-sequence queue heads
    }
    {
        Span sp(ctx, 1);
        for (int c = 0; c < NCLASS; c++) {
            if (!ncls[c]) continue;
            TraceParams T = Tbase;
            T.res = ctx->d_res.as<int2>();
            T.cand = ctx->d_cand_sorted[c].as<uint32_t>();
            T.cand_key = ctx->d_cand_key_sorted[c].as<uint16_t>();
            T.cand_rec = ctx->d_cand_rec[c].as<CandRec>();
            T.head = ctx->d_head.as<unsigned long long>();
            T.class_seqs = ctx->d_class_seqs[c].as<int32_t>();
            T.n_class_seqs = (int)ctx->class_seqs[c].size();
            T.ncand = ctx->d_ncand.as<unsigned long long>() + c;
            T.fb2 = ctx->d_fb2[c].as<uint32_t>(); T.fb2_cnt = ctx->d_fb2cnt.as<unsigned long long>();
            T.fallback = ctx->d_fallback[c].as<uint32_t>();
            T.nfallback = ctx->d_nfallback.as<unsigned long long>() + c;
            int rc = launch_trace_class(ctx, c, T, ncls[c], qmax[c]);
            if (rc) return rc;
            if (cnt) cnt->trace_launches += 3;
        }
    }
    return LSW_OK;
}

template <int NR>
int launch_ssw(lsw_ctx* ctx, SswParams P, int64_t n_class_tasks, const ScoreParams* K1) {
    if (K1) {
        // K1: the forward scan of every task of the class with packed DPX instructions (lsw_affine.cuh); it queues the tasks
        // that need the rest (reverse scan, banded traceback) as candidates
        const size_t smem = score_smem_bytes<NR>();
        static LaunchCfg cfg;
        int occ = 0;
        CK(cfg.get(ctx->device, k_score_affine<NR>, SCORE_WARPS * 32, smem, occ));
        if (occ < 1) return fail(ctx, LSW_E_INTERNAL, "affine score kernel does not fit on an SM");
        const int blocks1 = grid_for((n_class_tasks + 1) / 2, SCORE_WARPS * 32, ctx->num_sms * occ);
        k_score_affine<NR><<<blocks1, SCORE_WARPS * 32, smem, ctx->stream>>>(*K1);
        CK(cudaGetLastError());
    }
    // tier 1: one thread per task (or per candidate of K1), band <= SSW_BW1 in local memory
    const int blocks = grid_for(n_class_tasks, SSW_THREADS, ctx->num_sms * 16);
    k_ssw_round<NR, SSW_BW1><<<blocks, SSW_THREADS, 0, ctx->stream>>>(P, nullptr);
    CK(cudaGetLastError());
    // tier 2: the tasks whose band outgrew it, direction bytes in a global scratch
    const int blocks2 = 32;
    CK(ctx->d_scratch.ensure((size_t)blocks2 * SSW_THREADS * SswBand<SSW_BW2>::DIR_BYTES));
    SswParams P2 = P;
    P2.fb_in = P.fallback; P2.fb_in_cnt = P.nfallback;
    P2.fallback = nullptr; P2.nfallback = ctx->d_fb2cnt.as<unsigned long long>();     // (a tier-2 overflow is an error, not a list)
    k_ssw_round<NR, SSW_BW2><<<blocks2, SSW_THREADS, 0, ctx->stream>>>(P2, ctx->d_scratch.as<uint8_t>());
    CK(cudaGetLastError());
    return LSW_OK;
}

int launch_ssw_class(lsw_ctx* ctx, int c, const SswParams& P, int64_t n, const ScoreParams* K1) {
    switch (CLASS_NR[c]) {
        case 16: return launch_ssw<16>(ctx, P, n, K1);
        case 20: return launch_ssw<20>(ctx, P, n, K1);
        case 24: return launch_ssw<24>(ctx, P, n, K1);
        case 28: return launch_ssw<28>(ctx, P, n, K1);
        case 32: return launch_ssw<32>(ctx, P, n, K1);
        case 40: return launch_ssw<40>(ctx, P, n, K1);
        case 48: return launch_ssw<48>(ctx, P, n, K1);
        case 56: return launch_ssw<56>(ctx, P, n, K1);
        default: return launch_ssw<64>(ctx, P, n, K1);
    }
}

// One recursion round of the default (skbio / SSW) seeding path.  Per register-row class: K1 (packed forward scan, when the
// scores fit its 8.8 cells: match * len(primer) <= 127) and the per-task kernel that finishes the alignment (reverse scan,
// banded traceback, hit test, children) in two band tiers; without K1 the per-task kernel does the forward scan itself.
int run_round_ssw(lsw_ctx* ctx, const TaskBuf& cur, int segi, const std::vector<int64_t>& h_seg, int64_t n, SswParams P, lsw_counters* cnt) {
    CK(ctx->d_nfallback.ensure(NCLASS * 8));
    CK(cudaMemsetAsync(ctx->d_nfallback.p, 0, NCLASS * 8, ctx->stream));
    CK(ctx->d_fb2cnt.ensure((size_t)std::max(ctx->nseq, 1) * 8));
    CK(cudaMemsetAsync(ctx->d_fb2cnt.p, 0, 8, ctx->stream));
    P.codes = ctx->d_codes.as<uint8_t>(); P.roff = ctx->d_roff.as<int64_t>();
    P.t = cur.view();
    P.qcodes = ctx->d_qcodes.as<uint8_t>(); P.qlen = ctx->d_qlen.as<int32_t>();
    P.seg = ctx->d_seg[segi].as<int64_t>();
    P.match = ctx->ssw_match; P.mismatch = ctx->ssw_mismatch; P.gap_open = ctx->ssw_open; P.gap_extend = ctx->ssw_extend;
    P.counters = ctx->d_counters.as<unsigned long long>();
    P.errflag = ctx->d_err.as<int32_t>();
    bool packed = !ctx->env_ssw_scalar;
    for (int s = 0; s < ctx->nseq; s++) if (ctx->ssw_match * ctx->qlen[s] > 127) packed = false;
    ScoreParams K1;
    if (packed) {
        CK(ctx->d_res.ensure((size_t)n * sizeof(int2)));
        CK(ctx->d_head.ensure((size_t)ctx->nseq * 8));
        CK(ctx->d_ncand.ensure(NCLASS * 8));
        CK(cudaMemsetAsync(ctx->d_head.p, 0, (size_t)ctx->nseq * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_ncand.p, 0, NCLASS * 8, ctx->stream));
        CK(ctx->d_kt.ensure((size_t)n * sizeof(KTask)));
        k_make_ktasks<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(cur.view(), ctx->d_roff.as<int64_t>(), ctx->d_kt.as<KTask>(), n);
        CK(cudaGetLastError());
        if (cnt) cnt->other_launches++;
        fill_score_params(ctx, K1, cur, segi);
        K1.kt = ctx->d_kt.as<KTask>();
        K1.s_match = SC * ctx->ssw_match; K1.s_mis = SC * ctx->ssw_mismatch;
        K1.aff_open = ctx->ssw_open; K1.aff_extend = ctx->ssw_extend;
        K1.push_cands = 1; K1.count_ref = 1; K1.block_steps = BLOCK_STEPS;
        P.res = ctx->d_res.as<int2>();
    }
    Span sp(ctx, (P.mode == 0 && P.round == 0) ? 4 : 0);
    for (int c = 0; c < NCLASS; c++) {
        int64_t ncls = 0;
        for (int s : ctx->class_seqs[c]) ncls += h_seg[s + 1] - h_seg[s];
        if (!ncls) continue;
        CK(ctx->d_fallback[c].ensure((size_t)ncls * 4));
        P.class_seqs = ctx->d_class_seqs[c].as<int32_t>();
        P.n_class_seqs = (int)ctx->class_seqs[c].size();
        P.fallback = ctx->d_fallback[c].as<uint32_t>();
        P.nfallback = ctx->d_nfallback.as<unsigned long long>() + c;
        if (packed) {
            CK(ctx->d_cand[c].ensure((size_t)ncls * 4)); CK(ctx->d_cand_key[c].ensure((size_t)ncls * 2));
            CK(ctx->d_cand_rec[c].ensure((size_t)ncls * sizeof(CandRec)));
            K1.class_seqs = P.class_seqs; K1.n_class_seqs = P.n_class_seqs;
            K1.cand = ctx->d_cand[c].as<uint32_t>(); K1.cand_key = ctx->d_cand_key[c].as<uint16_t>();
            K1.cand_rec = ctx->d_cand_rec[c].as<CandRec>();
            K1.ncand = ctx->d_ncand.as<unsigned long long>() + c;
            P.cand_rec = K1.cand_rec; P.ncand = K1.ncand;
        }
        if (int rc = launch_ssw_class(ctx, c, P, ncls, packed ? &K1 : nullptr)) return rc;
        if (cnt) { cnt->score_launches += packed ? 1 : 0; cnt->trace_launches += 2; }
    }
    return LSW_OK;
}

void collect_times(lsw_ctx* ctx, lsw_counters* cnt) {
    for (const Timed& t : ctx->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) != cudaSuccess) continue;
        if (t.kind == 0) cnt->ms_score += ms;
        else if (t.kind == 4) { cnt->ms_score += ms; cnt->ms_score_round0 += ms; }
        else if (t.kind == 1) cnt->ms_trace += ms;
        else if (t.kind == 2) cnt->ms_other += ms;
        else cnt->ms_total += ms;
    }
    ctx->timed.clear();
    ctx->ev_used = 0;
}

}  // namespace

// ================================ C ABI ====================================================

extern "C" {

int lsw_version(void) { return 1; }

const char* lsw_last_error(const lsw_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int lsw_create(int device, lsw_ctx** out) {
    if (!out) { g_create_error = "lsw_create: out is NULL"; return LSW_E_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_create_error = std::string("lsw_create: no CUDA device (") + cudaGetErrorString(e) +
                         "); this engine has no CPU fallback";
        return LSW_E_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "lsw_create: device index out of range"; return LSW_E_ARG; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return LSW_E_CUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return LSW_E_CUDA; }
    if (prop.major < 10) {
        g_create_error = "lsw_create: kernels are built for sm_100a (B200) only";
        return LSW_E_CUDA;
    }
    lsw_ctx* ctx = new lsw_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete ctx; return LSW_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    read_env(ctx);
    *out = ctx;
    return LSW_OK;
}

void lsw_destroy(lsw_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf* b : all_bufs(ctx)) b->release();
    for (cudaEvent_t ev : ctx->ev_pool) cudaEventDestroy(ev);
    for (int k = 0; k < 2; k++) { if (ctx->ev_copied[k]) cudaEventDestroy(ctx->ev_copied[k]); if (ctx->ev_coded[k]) cudaEventDestroy(ctx->ev_coded[k]); }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int lsw_set_stream(lsw_ctx* ctx, void* cuda_stream) {
    if (!ctx) return LSW_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return LSW_OK;
}

int lsw_set_read_base(lsw_ctx* ctx, int64_t base) {
    if (!ctx || base < 0 || base >= (1ll << 31)) return LSW_E_ARG;
    ctx->read_base = (int32_t)base;
    return LSW_OK;
}

int lsw_set_overlap(lsw_ctx* ctx, int allowed_overlap) {
    if (!ctx) return LSW_E_ARG;
    ctx->allowed_overlap = allowed_overlap < 0 ? -1 : allowed_overlap;
    return LSW_OK;
}

int lsw_set_hit_capacity(lsw_ctx* ctx, int64_t cap) {
    if (!ctx || cap < 0) return LSW_E_ARG;
    ctx->hit_cap_user = cap;
    return LSW_OK;
}

int lsw_set_scoring(lsw_ctx* ctx, int match, int mismatch, int gap, int gap_ext) {
    if (!ctx) return LSW_E_ARG;
    if (gap != gap_ext)
        return fail(ctx, LSW_E_UNSUPPORTED, "gap_penalty != gap_extension_penalty is not implemented (LAMPrey uses -1/-1)");
    // mismatch must be strictly negative: K1 lets a lane run a few dead columns past a slice end and
    // relies on every move through them costing at least 1, so that they can never tie the arg-max.
    if (match <= 0 || mismatch >= 0 || gap >= 0)
        return fail(ctx, LSW_E_UNSUPPORTED, "need match > 0, mismatch < 0, gap < 0");
    if (match * LSW_MAX_QUERY_LEN - gap > 127 && match * 1 - gap > 127)
        return fail(ctx, LSW_E_UNSUPPORTED, "scores do not fit the packed 8.8 fixed-point cells");
    // the profile entry of a mismatch, 256 * (mismatch + |gap|), is one s16 half (lsw_score.cuh build_profile): below -32768 it
    // would wrap and a mismatch would score positive
    if (mismatch - gap < -128)
        return fail(ctx, LSW_E_UNSUPPORTED, "mismatch + |gap| must be >= -128 for the packed 8.8 fixed-point cells");
    ctx->match = match; ctx->mismatch = mismatch; ctx->g = -gap;
    ctx->ssw = false;
    return LSW_OK;          // scoring and queries may be set in either order: the pair is checked when work is submitted
}

int lsw_set_ssw(lsw_ctx* ctx, int match, int mismatch, int gap_open, int gap_extend) {
    if (!ctx) return LSW_E_ARG;
    if (match <= 0 || match > 100 || mismatch > 0 || mismatch < -100 || gap_open <= 0 || gap_open > 100 || gap_extend <= 0 ||
        gap_extend > gap_open)
        return fail(ctx, LSW_E_UNSUPPORTED, "SSW scoring: need 0 < match <= 100, -100 <= mismatch <= 0, 0 < gap_extend <= gap_open <= 100");
    ctx->ssw_match = match; ctx->ssw_mismatch = mismatch; ctx->ssw_open = gap_open; ctx->ssw_extend = gap_extend;
    ctx->ssw = true;
    return LSW_OK;
}

// Scores are kept as 8.8 fixed point in 16-bit halves, biased by g: the largest value a cell can hold must stay <= 127.
static int check_scoring_fits(lsw_ctx* ctx) {
    if (ctx->ssw) return LSW_OK;          // scalar int32 cells: lsw_set_ssw's bounds are enough
    for (int s = 0; s < ctx->nseq; s++)
        if (ctx->match * ctx->qlen[s] + ctx->g > 127)
            return fail(ctx, LSW_E_UNSUPPORTED, "match * len(query) + |gap| must be <= 127 for the packed cells");
    return LSW_OK;
}

int lsw_set_queries(lsw_ctx* ctx, int n, const uint8_t* concat, const int32_t* offs) {
    if (!ctx || n <= 0 || n > 255 || !concat || !offs) return ctx ? fail(ctx, LSW_E_ARG, "lsw_set_queries: bad argument (1..255 sequences)") : LSW_E_ARG;
    CK(cudaSetDevice(ctx->device));
    std::vector<int32_t> qlen(n), class_of(n);
    std::vector<uint8_t> qcodes((size_t)n * QSTRIDE, 0xFF);
    for (int s = 0; s < n; s++) {
        const int len = offs[s + 1] - offs[s];
        if (len < 1 || len > LSW_MAX_QUERY_LEN)
            return fail(ctx, LSW_E_UNSUPPORTED, "query length must be 1..63 (all rows are kept in registers)");
        qlen[s] = len;
        for (int j = 0; j < len; j++) {
            const int code = encode_base(concat[offs[s] + j]) < CODE_OTHER ? (int)encode_base(concat[offs[s] + j]) : -1;
            if (code < 0)
                return fail(ctx, LSW_E_UNSUPPORTED, "queries must be ACGT only (swalign would match a non-ACGT query character against the same character in a read)");
            qcodes[(size_t)s * QSTRIDE + j] = (uint8_t)code;
        }
        int c = 0;
        while (CLASS_NR[c] < len) c++;
        class_of[s] = c;
    }
    ctx->nseq = n; ctx->qlen = qlen; ctx->qcodes = qcodes; ctx->class_of = class_of;
    for (int c = 0; c < NCLASS; c++) ctx->class_seqs[c].clear();
    for (int s = 0; s < n; s++) ctx->class_seqs[class_of[s]].push_back(s);
    CK(ctx->d_qcodes.ensure(qcodes.size()));
    CK(ctx->d_qlen.ensure((size_t)n * 4));
    CK(cudaMemcpyAsync(ctx->d_qcodes.p, qcodes.data(), qcodes.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_qlen.p, qlen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_class_of.ensure((size_t)n * 4));
    CK(cudaMemcpyAsync(ctx->d_class_of.p, class_of.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    for (int c = 0; c < NCLASS; c++) {
        if (ctx->class_seqs[c].empty()) continue;
        CK(ctx->d_class_seqs[c].ensure(ctx->class_seqs[c].size() * 4));
        CK(cudaMemcpyAsync(ctx->d_class_seqs[c].p, ctx->class_seqs[c].data(), ctx->class_seqs[c].size() * 4,
                           cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return LSW_OK;
}

int lsw_upload_reads(lsw_ctx* ctx, int64_t n, const uint8_t* ascii, const int64_t* offs) {
    if (!ctx || n < 0 || !offs || (n > 0 && !ascii && offs[n] > offs[0])) return ctx ? fail(ctx, LSW_E_ARG, "lsw_upload_reads: bad argument") : LSW_E_ARG;
    if (n >= (1ll << 26)) return fail(ctx, LSW_E_ARG, "at most 2^26 - 1 reads per batch");
    CK(cudaSetDevice(ctx->device));
    // a failed upload leaves the context EMPTY (n_reads = 0), never describing the previous batch with the new lengths
    ctx->n_reads = 0; ctx->total_bases = 0; ctx->n_hits = 0; ctx->hits_pruned = false;
    // one pass over the offsets: lengths, padded size, and the block-best rows of lsw_reuse.cuh (BB_BLOCK columns per block,
    // BB_GROUP blocks per second-level entry)
    ctx->h_rlen.resize((size_t)n);
    ctx->h_bb_off.resize((size_t)n + 1); ctx->h_bb1_off.resize((size_t)n + 1);
    ctx->h_bb_off[0] = 0; ctx->h_bb1_off[0] = 0;
    int64_t padded_total = 0;
    for (int64_t i = 0; i < n; i++) {
        const int64_t len = offs[i + 1] - offs[i];
        if (len < 0 || len >= (1ll << 27)) { ctx->h_rlen.clear(); return fail(ctx, LSW_E_ARG, "read length must be in [0, 2^27)"); }
        ctx->h_rlen[i] = (int32_t)len;
        padded_total += (len + 15) & ~15ll;
        const int64_t nb = (len + BB_BLOCK - 1) / BB_BLOCK;
        ctx->h_bb_off[i + 1] = ctx->h_bb_off[i] + nb;
        ctx->h_bb1_off[i + 1] = ctx->h_bb1_off[i] + (nb + BB_GROUP - 1) / BB_GROUP;
    }
    const int64_t total_bases = n ? offs[n] - offs[0] : 0;
    ctx->total_bases = total_bases;
    ctx->bb_total = ctx->h_bb_off[n];
    ctx->bb1_total = ctx->h_bb1_off[n];
    CK(ctx->d_bb_off.ensure((size_t)(n + 1) * 8));
    CK(cudaMemcpyAsync(ctx->d_bb_off.p, ctx->h_bb_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_bb1_off.ensure((size_t)(n + 1) * 8));
    CK(cudaMemcpyAsync(ctx->d_bb1_off.p, ctx->h_bb1_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    // a slice must span more than two groups (2048 columns) before the second level saves lookups
    ctx->use_bb1 = n > 0 && total_bases / n >= 4 * BB_GROUP * BB_BLOCK;
    if (const char* e = getenv("LSW_BB_L1")) ctx->use_bb1 = atoi(e) != 0;
    // 256 B front pad (K2 reads a few codes before a window start, masked) and 256 B tail pad (K1 reads
    // whole 64-byte blocks past a slice end)
    ctx->codes_bytes = 256 + padded_total + 256;
    CK(ctx->d_offs_in.ensure((size_t)(n + 1) * 8));
    CK(ctx->d_roff.ensure((size_t)(n + 1) * 8));
    CK(ctx->d_rlen.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    CK(ctx->d_codes.ensure((size_t)ctx->codes_bytes));
    CK(cudaMemcpyAsync(ctx->d_offs_in.p, offs, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));   // absolute; the kernel subtracts the chunk's base
    if (n) CK(cudaMemcpyAsync(ctx->d_rlen.p, ctx->h_rlen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_codes.as<uint8_t>() + 256 + padded_total, CODE_OTHER, 256, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_codes.as<uint8_t>(), CODE_OTHER, 256, ctx->stream));
    if (n) {
        // The device copy of the reads is laid out LONGEST READ FIRST: round 0 queues the reads in that order
        // (the few very long reads of an ONT length distribution must not form the tail of a launch), and
        // consecutive tasks then touch consecutive memory.  Read ids stay the caller's: roff[] is indexed by
        // the original id and only its VALUES follow the sorted layout.
        CK(ctx->d_order.ensure((size_t)n * 4)); CK(ctx->d_order_tmp.ensure((size_t)n * 4));
        CK(ctx->d_len_tmp[0].ensure((size_t)n * 4)); CK(ctx->d_len_tmp[1].ensure((size_t)n * 4));
        CK(ctx->d_pos.ensure((size_t)(n + 1) * 8));
        k_iota<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(ctx->d_order_tmp.as<int32_t>(), n);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->d_len_tmp[0].p, ctx->d_rlen.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, ctx->d_len_tmp[0].as<int32_t>(), ctx->d_len_tmp[1].as<int32_t>(),
                                                     ctx->d_order_tmp.as<int32_t>(), ctx->d_order.as<int32_t>(), (int)n, 0, 27, ctx->stream));
        CK(ctx->d_cubtemp.ensure(tmp));
        CK(cub::DeviceRadixSort::SortPairsDescending(ctx->d_cubtemp.p, tmp, ctx->d_len_tmp[0].as<int32_t>(), ctx->d_len_tmp[1].as<int32_t>(),
                                                     ctx->d_order_tmp.as<int32_t>(), ctx->d_order.as<int32_t>(), (int)n, 0, 27, ctx->stream));
        // offsets in the sorted layout: exclusive scan of the 16-byte padded sorted lengths, scattered to roff[order[i]]
        auto padded = thrust::make_transform_iterator((const int32_t*)ctx->d_len_tmp[1].as<int32_t>(), PadLen());
        tmp = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, padded, ctx->d_pos.as<long long>(), (int)n, ctx->stream));
        CK(ctx->d_cubtemp.ensure(tmp));
        CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtemp.p, tmp, padded, ctx->d_pos.as<long long>(), (int)n, ctx->stream));
        k_scatter_offsets<<<grid_for(n, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            ctx->d_order.as<int32_t>(), ctx->d_pos.as<long long>(), ctx->d_roff.as<int64_t>(), n);
        CK(cudaGetLastError());
        // The bases stream through two staging buffers of UPLOAD_CHUNK bytes: the copy stream brings chunk k + 1 in while chunk
        // k is re-coded into its place in the sorted layout, so the ASCII text never exists on the device as a whole (it used to:
        // a second buffer the size of the batch).  Chunks end on read boundaries; a read longer than a chunk is a chunk of its own.
        int64_t max_chunk = 0;
        std::vector<int64_t> cuts(1, 0);
        while (cuts.back() < n) {
            const int64_t r0 = cuts.back();
            int64_t r1 = std::upper_bound(offs + r0, offs + n + 1, offs[r0] + UPLOAD_CHUNK) - offs - 1;
            if (r1 <= r0) r1 = r0 + 1;
            cuts.push_back(r1);
            max_chunk = std::max(max_chunk, offs[r1] - offs[r0]);
        }
        const int nchunks = (int)cuts.size() - 1;
        const int nbuf = nchunks > 1 ? 2 : 1;
        CK(ctx->d_ascii.ensure((size_t)std::max<int64_t>(max_chunk, 1) * nbuf));
        if (!ctx->copy_stream) {
            CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            for (int k = 0; k < 2; k++) { CK(cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_coded[k], cudaEventDisableTiming)); }
        }
        CK(cudaEventRecord(ctx->ev_coded[0], ctx->stream));           // the offsets are in place (and the staging buffers are free)
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_coded[0], 0));
        for (int c = 0; c < nchunks; c++) {
            const int64_t r0 = cuts[c], r1 = cuts[c + 1], bytes = offs[r1] - offs[r0];
            uint8_t* stage = ctx->d_ascii.as<uint8_t>() + (size_t)(c & 1) * (size_t)max_chunk;
            if (c >= 2) CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_coded[c & 1], 0));     // chunk c - 2 has left this buffer
            if (bytes) CK(cudaMemcpyAsync(stage, ascii + offs[r0], (size_t)bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_copied[c & 1], ctx->copy_stream));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[c & 1], 0));
            const int blocks = grid_for((r1 - r0) * 32, 256, ctx->num_sms * 16);
            k_encode_reads<<<blocks, 256, 0, ctx->stream>>>(stage, ctx->d_offs_in.as<int64_t>(), offs[r0], ctx->d_roff.as<int64_t>(),
                                                            ctx->d_codes.as<uint8_t>(), r0, r1);
            CK(cudaGetLastError());
            CK(cudaEventRecord(ctx->ev_coded[c & 1], ctx->stream));
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));      // the caller's buffers may go away
    ctx->n_reads = n;                            // committed only now: every step above has succeeded
    return LSW_OK;
}

int lsw_find_all(lsw_ctx* ctx, double thr, int64_t* n_hits, lsw_counters* counters) {
    if (!ctx) return LSW_E_ARG;
    if (!(thr > 0.0)) return fail(ctx, LSW_E_ARG, "identity_threshold must be > 0 (the reference recurses forever otherwise)");
    if (ctx->nseq <= 0) return fail(ctx, LSW_E_STATE, "lsw_set_queries has not been called");
    if (int rc = check_scoring_fits(ctx)) return rc;
    CK(cudaSetDevice(ctx->device));
    read_env(ctx);
    lsw_counters cnt;
    memset(&cnt, 0, sizeof cnt);
    ctx->timed.clear(); ctx->ev_used = 0;
    ctx->n_hits = 0;
    if (n_hits) *n_hits = 0;
    const int64_t n0 = ctx->n_reads * ctx->nseq;
    if (n0 >= (1ll << 30)) return fail(ctx, LSW_E_ARG, "reads x sequences must stay below 2^30 tasks per batch");
    if (n0 == 0) { if (counters) *counters = cnt; return LSW_OK; }

    std::vector<uint8_t> tab;
    if (ctx->ssw) {      // K1 of the skbio path queues every task with a positive score (the per-task kernel applies the hit test)
        tab.assign((size_t)ctx->nseq * 128, 1);
        for (int s = 0; s < ctx->nseq; s++) tab[(size_t)s * 128] = 0;
    } else
    build_cand_ok(ctx->nseq, ctx->qlen.data(), ctx->match, ctx->mismatch, ctx->g, thr, false, tab);
    CK(ctx->d_cand_ok.ensure(tab.size()));
    CK(cudaMemcpyAsync(ctx->d_cand_ok.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    // head / warm-up length of child slices (lsw_reuse.cuh): only cells that could pass the hit test must be the whole read's cells
    std::vector<int32_t> wx;
    build_reuse_wx(ctx->nseq, ctx->qlen.data(), ctx->match, ctx->g, tab, ctx->env_full_head || ctx->ssw, wx);
    CK(ctx->d_wx.ensure(wx.size() * sizeof(int32_t)));
    CK(cudaMemcpyAsync(ctx->d_wx.p, wx.data(), wx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));

    int64_t hit_cap = ctx->hit_cap_user > 0 ? ctx->hit_cap_user : ctx->total_bases / 8 + 4096;
    CK(ctx->d_hits.ensure((size_t)hit_cap * sizeof(lsw_hit)));
    CK(ctx->d_nhits.ensure(8));
    CK(ctx->d_counters.ensure(32));
    CK(ctx->d_err.ensure(4));
    CK(cudaMemsetAsync(ctx->d_nhits.p, 0, 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 32, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_err.p, 0, 4, ctx->stream));
    CK(ctx->d_seg[0].ensure((size_t)(ctx->nseq + 1) * 8));
    CK(ctx->d_seg[1].ensure((size_t)(ctx->nseq + 1) * 8));

    Timed tt; tt.kind = 3; tt.e0 = next_event(ctx); tt.e1 = next_event(ctx);
    if (tt.e0) cudaEventRecord(tt.e0, ctx->stream);
    std::vector<int64_t> h_seg((size_t)ctx->nseq + 1);
    for (int s = 0; s <= ctx->nseq; s++) h_seg[s] = (int64_t)s * ctx->n_reads;
    CK(cudaMemcpyAsync(ctx->d_seg[0].p, h_seg.data(), h_seg.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    int cur = 0;
    CK(ctx->tb[0].ensure((size_t)n0, false));
    {
        Span sp(ctx, 2);
        k_init_tasks<<<grid_for(n0, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            ctx->tb[0].view(), ctx->d_rlen.as<int32_t>(), ctx->d_order.as<int32_t>(), ctx->n_reads, ctx->nseq);
        CK(cudaGetLastError());
        cnt.other_launches++;
    }

    // block bests of round 0 for the child tasks: 4 B per BB_BLOCK columns per (read, sequence)
    bool reuse = ctx->reuse && !ctx->env_no_reuse && !ctx->ssw;    // (SSW's "first maximum wins" end rule has no block-best merge yet)
    // the block bests are an optimisation (4 B per 32 columns per sequence): without room for them every round simply
    // aligns whole slices -- same results, more cells
    if (reuse && ctx->d_blockbest.ensure((size_t)std::max<int64_t>(1, ctx->bb_total) * ctx->nseq * 4) != cudaSuccess) {
        (void)cudaGetLastError();
        reuse = false;
        cnt.reuse_dropped = 1;          // reported, not silent: lsw_counters.reuse_dropped (and a line on stderr)
        fprintf(stderr, "[lsw] warning: no room for the round-0 block bests (%.1f GB): child rounds align whole slices\n",
                (double)ctx->bb_total * ctx->nseq * 4 / 1e9);
    }
    int64_t n = n0;
    int round = 0;
    size_t dbg_from = 0;
    int rc = LSW_OK;
    unsigned long long h_nhits = 0;
    while (n > 0) {
        // the scans, the 32-bit child positions and K1's task indices are int32: a round may hold at most 2^30 - 1 tasks
        if (n >= (1ll << 30)) { rc = fail(ctx, LSW_E_ARG, "a recursion round exceeds 2^30 tasks: split the batch"); break; }
        const int64_t n2 = 2 * n;
        rc = ctx->child.ensure((size_t)n2, false) == cudaSuccess ? LSW_OK : LSW_E_CUDA;
        if (rc) { ctx->err = "out of device memory (child tasks)"; break; }
        CK(ctx->d_child_flag.ensure((size_t)n2 + 1));
        CK(ctx->d_pos.ensure((size_t)(n2 + 1) * 4));
        CK(cudaMemsetAsync(ctx->d_child_flag.p, 0, (size_t)n2 + 1, ctx->stream));

        TraceParams T;
        fill_trace_params(ctx, T, ctx->tb[cur]);
        T.mode = 0; T.thr = thr; T.round = round;
        T.hits = ctx->d_hits.p; T.hits_cap = hit_cap; T.nhits = ctx->d_nhits.as<unsigned long long>();
        T.child_flag = ctx->d_child_flag.as<uint8_t>();
        T.child = ctx->child.view();
        if (ctx->ssw) {
            SswParams S;
            memset(&S, 0, sizeof S);
            S.mode = 0; S.thr = thr; S.round = round;
            S.hits = T.hits; S.hits_cap = T.hits_cap; S.nhits = T.nhits;
            S.child_flag = T.child_flag; S.child = T.child;
            rc = run_round_ssw(ctx, ctx->tb[cur], cur, h_seg, n, S, &cnt);
        } else {
            // (the reuse path stays on in the late, tiny rounds as well: it also bounds the SERIAL work of a lane -- a 10 kb child
            // slice is a 100-column head and tail there, but milliseconds of one lane's time when aligned whole; measured)
            rc = run_round(ctx, ctx->tb[cur], cur, h_seg, n, T, &cnt, reuse ? (round == 0 ? 1 : 2) : 0);
        }
        if (rc) break;
        if (reuse && round == 0 && ctx->use_bb1 && ctx->bb1_total > 0) {
            Span sp(ctx, 2);
            CK(ctx->d_bb1.ensure((size_t)ctx->bb1_total * ctx->nseq * 4));
            k_build_bb1<<<grid_for(ctx->bb1_total * ctx->nseq, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
                ctx->d_blockbest.as<uint32_t>(), ctx->d_bb_off.as<int64_t>(), ctx->bb_total, ctx->d_bb1.as<uint32_t>(),
                ctx->d_bb1_off.as<int64_t>(), ctx->bb1_total, ctx->n_reads, ctx->nseq, ctx->g);
            CK(cudaGetLastError());
            cnt.other_launches++;
        }

        {
            Span sp(ctx, 2);
            auto it = thrust::make_transform_iterator((const uint8_t*)ctx->d_child_flag.as<uint8_t>(), FlagToInt());
            size_t tmp = 0;
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, ctx->d_pos.as<int32_t>(), (int)(n2 + 1), ctx->stream));
            CK(ctx->d_cubtemp.ensure(tmp));
            CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtemp.p, tmp, it, ctx->d_pos.as<int32_t>(), (int)(n2 + 1), ctx->stream));
            cnt.other_launches += 2;
        }
        // total children (pos[n2]) decides whether the preallocated next list is large enough
        int32_t total_next = 0;
        CK(cudaMemcpyAsync(&total_next, ctx->d_pos.as<int32_t>() + n2, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(ctx->tb[cur ^ 1].ensure((size_t)std::max<int32_t>(total_next, 1), false));
        {
            Span sp(ctx, 2);
            k_scatter_children<<<grid_for(n2, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
                ctx->child.view(), ctx->d_child_flag.as<uint8_t>(), ctx->d_pos.as<int32_t>(), n2,
                ctx->tb[cur ^ 1].view(), ctx->d_seg[cur].as<int64_t>(), ctx->d_seg[cur ^ 1].as<int64_t>(), ctx->nseq);
            CK(cudaGetLastError());
            cnt.other_launches++;
        }
        CK(cudaMemcpyAsync(h_seg.data(), ctx->d_seg[cur ^ 1].p, h_seg.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&h_nhits, ctx->d_nhits.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx->d_ncand.ensure(NCLASS * 8)); CK(ctx->d_nfallback.ensure(NCLASS * 8));
        if (ctx->ssw) CK(cudaMemsetAsync(ctx->d_ncand.p, 0, NCLASS * 8, ctx->stream));
        unsigned long long h_nc[NCLASS], h_nf[NCLASS], h_comp0 = 0;
        if (round == 0) CK(cudaMemcpyAsync(&h_comp0, ctx->d_counters.as<unsigned long long>() + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_nc, ctx->d_ncand.p, NCLASS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_nf, ctx->d_nfallback.p, NCLASS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int c = 0; c < NCLASS; c++) { cnt.traced += (int64_t)h_nc[c]; cnt.fallbacks += (int64_t)h_nf[c]; }
        if (ctx->env_debug) {
            float ms[3] = {0.f, 0.f, 0.f};
            for (size_t k = dbg_from; k < ctx->timed.size(); k++) {
                float m = 0.f;
                const int kd = ctx->timed[k].kind == 4 ? 0 : ctx->timed[k].kind;
                if (kd < 3 && cudaEventElapsedTime(&m, ctx->timed[k].e0, ctx->timed[k].e1) == cudaSuccess) ms[kd] += m;
            }
            dbg_from = ctx->timed.size();
            unsigned long long c4[4] = {0, 0, 0, 0};
            cudaMemcpy(c4, ctx->d_counters.p, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[lsw] round %d: tasks %lld -> children %d, hits so far %llu, score %.3f ms, trace %.3f ms, other %.3f ms; "
                            "cumulative lane-steps %llu, computed cells %llu, reference cells %llu\n",
                    round, (long long)n, total_next, h_nhits, ms[0], ms[1], ms[2], c4[2], c4[3], c4[1]);
        }
        if (round == 0) { cnt.top_tasks = n; cnt.computed_cells_round0 = (int64_t)h_comp0; }
        cur ^= 1;
        n = total_next;
        round++;
        if (round > 30000) { rc = fail(ctx, LSW_E_INTERNAL, "recursion did not terminate"); break; }
    }
    if (rc) return rc;
    cnt.rounds = round;

    // work counters and the internal-invariant flag
    unsigned long long h_cnt[4] = {0, 0, 0, 0};
    int32_t h_err = 0;
    CK(cudaMemcpyAsync(h_cnt, ctx->d_counters.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&h_err, ctx->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cnt.tasks = (int64_t)h_cnt[0]; cnt.cells = (int64_t)h_cnt[1]; cnt.steps = (int64_t)h_cnt[2]; cnt.computed_cells = (int64_t)h_cnt[3];
    cnt.top_cells = 0;
    {
        int64_t sumq = 0;
        for (int s = 0; s < ctx->nseq; s++) sumq += ctx->qlen[s];
        cnt.top_cells = sumq * ctx->total_bases;
    }
    cnt.hits = (int64_t)h_nhits;
    if (h_err) return fail(ctx, LSW_E_INTERNAL, "traceback window bound violated");
    if ((int64_t)h_nhits > hit_cap) {
        ctx->n_hits = 0;
        if (n_hits) *n_hits = (int64_t)h_nhits;
        if (counters) *counters = cnt;
        return fail(ctx, LSW_E_OVERFLOW, "hit buffer too small; call lsw_set_hit_capacity(*n_hits) and retry");
    }

    // order hits by (read, start, query)
    const int64_t nh = (int64_t)h_nhits;
    if (nh > 0) {
        Span sp(ctx, 2);
        CK(ctx->d_keys[0].ensure((size_t)nh * 8)); CK(ctx->d_keys[1].ensure((size_t)nh * 8));
        CK(ctx->d_vals[0].ensure((size_t)nh * 4)); CK(ctx->d_vals[1].ensure((size_t)nh * 4));
        CK(ctx->d_hits_sorted.ensure((size_t)nh * sizeof(lsw_hit)));
        auto bits_for = [](int64_t v) { int b = 1; while ((1ll << b) <= v) b++; return b; };      // bits that hold 0..v
        int max_len = 1;
        for (int32_t l : ctx->h_rlen) max_len = std::max(max_len, (int)l);
        const int qbits = bits_for(ctx->nseq - 1), sbits = bits_for(max_len), rbits = bits_for(ctx->n_reads - 1);
        const int key_bits = qbits + sbits + rbits;
        if (key_bits > 64) return fail(ctx, LSW_E_ARG, "batch too large for the 64-bit hit sort key (reads x longest read x sequences)");
        k_hit_keys<<<grid_for(nh, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            ctx->d_hits.as<lsw_hit>(), nh, ctx->d_keys[0].as<unsigned long long>(), ctx->d_vals[0].as<uint32_t>(), qbits, sbits);
        CK(cudaGetLastError());
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->d_keys[0].as<unsigned long long>(), ctx->d_keys[1].as<unsigned long long>(),
                                           ctx->d_vals[0].as<uint32_t>(), ctx->d_vals[1].as<uint32_t>(), (int)nh, 0, key_bits, ctx->stream));
        CK(ctx->d_cubtemp.ensure(tmp));
        CK(cub::DeviceRadixSort::SortPairs(ctx->d_cubtemp.p, tmp, ctx->d_keys[0].as<unsigned long long>(), ctx->d_keys[1].as<unsigned long long>(),
                                           ctx->d_vals[0].as<uint32_t>(), ctx->d_vals[1].as<uint32_t>(), (int)nh, 0, key_bits, ctx->stream));
        k_gather_hits<<<grid_for(nh, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
            ctx->d_hits.as<lsw_hit>(), ctx->d_vals[1].as<uint32_t>(), nh, ctx->d_hits_sorted.as<lsw_hit>(), ctx->read_base);
        CK(cudaGetLastError());
        cnt.other_launches += 3;
        if (ctx->allowed_overlap >= 0) {
            k_prune_overlaps<<<grid_for(ctx->n_reads, 128, ctx->num_sms * 8), 128, 0, ctx->stream>>>(
                ctx->d_hits_sorted.as<lsw_hit>(), nh, ctx->n_reads, ctx->read_base, ctx->allowed_overlap,
                ctx->ssw ? ctx->d_qlen.as<int32_t>() : nullptr);
            CK(cudaGetLastError());
            cnt.other_launches += 1;
        }
    }
    if (tt.e1) { cudaEventRecord(tt.e1, ctx->stream); ctx->timed.push_back(tt); }
    CK(cudaStreamSynchronize(ctx->stream));
    collect_times(ctx, &cnt);
    cnt.device_bytes = device_bytes(ctx);
    ctx->n_hits = nh;
    ctx->hits_pruned = nh > 0 && ctx->allowed_overlap >= 0;
    if (n_hits) *n_hits = nh;
    if (counters) *counters = cnt;
    return LSW_OK;
}

int lsw_ssw_subreads(lsw_ctx* ctx, int64_t n, const lsw_pair* pairs, const uint8_t* target, int32_t target_len, lsw_result* out,
                     uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used) {
    if (!ctx) return LSW_E_ARG;
    if (cigar_used) *cigar_used = 0;
    if (n < 0 || (n > 0 && (!pairs || !out)) || target_len < 0 || (target_len > 0 && !target))
        return fail(ctx, LSW_E_ARG, "lsw_ssw_subreads: bad argument");
    if (n == 0) return LSW_OK;
    if (target_len > 30000) return fail(ctx, LSW_E_UNSUPPORTED, "lsw_ssw_subreads: target reference longer than 30000 bases");
    int maxq = 1;
    for (int64_t i = 0; i < n; i++) {
        const lsw_pair& p = pairs[i];
        if (p.read < 0 || p.read >= ctx->n_reads || p.start < 0 || p.end < p.start || p.end > ctx->h_rlen[p.read])
            return fail(ctx, LSW_E_ARG, "lsw_ssw_subreads: sub-read out of range");
        maxq = std::max(maxq, p.end - p.start);
    }
    CK(cudaSetDevice(ctx->device));
    // banded_sw's direction bytes: (2 band + 1) x query span per sub-read.  An aligned query span cannot exceed the target by
    // more than the insertions its score pays for, so 4 x target rows at a band of 64 covers what real data produces; beyond that
    // the call reports LSW_E_UNSUPPORTED, never a guess.  One thread per sub-read in flight: as many as ~2 GB of scratch allow.
    const long long rows_cap = std::min<long long>(maxq, 4ll * target_len + 64);
    const long long dir_cap = std::min<long long>(std::max<long long>(129 * rows_cap, 16 << 10), 1 << 20);
    const long long per_slot = dir_cap + 5ll * maxq;
    long long slots = std::min<long long>(32768, (2ll << 30) / per_slot);
    slots = std::max<long long>(128, std::min<long long>(slots, ((n + 127) / 128) * 128));
    const int nslots = (int)((slots / 128) * 128);
    DevBuf d_pairs, d_target, d_he, d_qbuf, d_dir;
    auto release = [&]() { d_pairs.release(); d_target.release(); d_he.release(); d_qbuf.release(); d_dir.release(); };
    std::vector<uint8_t> tcodes((size_t)std::max(target_len, 1));
    for (int i = 0; i < target_len; i++) tcodes[i] = encode_base(target[i]);
    cudaError_t e = cudaSuccess;
    if ((e = d_pairs.ensure((size_t)n * sizeof(lsw_pair))) != cudaSuccess || (e = d_target.ensure(tcodes.size())) != cudaSuccess ||
        (e = d_he.ensure((size_t)maxq * nslots * 4)) != cudaSuccess || (e = d_qbuf.ensure((size_t)maxq * nslots)) != cudaSuccess ||
        (e = d_dir.ensure((size_t)dir_cap * nslots)) != cudaSuccess) {
        release();
        ctx->err = std::string("lsw_ssw_subreads: ") + cudaGetErrorString(e);
        return LSW_E_CUDA;
    }
    int rc = LSW_OK;
    unsigned long long h_ncig = 0;
    int32_t h_err = 0;
    do {
#define CKB(call) if ((e = (call)) != cudaSuccess) { ctx->err = std::string("lsw_ssw_subreads: ") + cudaGetErrorString(e); rc = LSW_E_CUDA; break; }
        CKB(ctx->d_results.ensure((size_t)n * sizeof(lsw_result)));
        CKB(ctx->d_err.ensure(4)); CKB(ctx->d_ncigar.ensure(8));
        const bool want_cigar = cigar_buf != nullptr && cigar_cap > 0;
        if (want_cigar) CKB(ctx->d_cigar.ensure((size_t)cigar_cap * 4));
        CKB(cudaMemsetAsync(ctx->d_err.p, 0, 4, ctx->stream));
        CKB(cudaMemsetAsync(ctx->d_ncigar.p, 0, 8, ctx->stream));
        CKB(cudaMemcpyAsync(d_pairs.p, pairs, (size_t)n * sizeof(lsw_pair), cudaMemcpyHostToDevice, ctx->stream));
        CKB(cudaMemcpyAsync(d_target.p, tcodes.data(), tcodes.size(), cudaMemcpyHostToDevice, ctx->stream));
        SswGenParams G;
        memset(&G, 0, sizeof G);
        G.codes = ctx->d_codes.as<uint8_t>(); G.roff = ctx->d_roff.as<int64_t>();
        G.pairs = d_pairs.as<lsw_pair>(); G.n = n;
        G.tcodes = d_target.as<uint8_t>(); G.tlen = target_len;
        G.match = ctx->ssw_match; G.mismatch = ctx->ssw_mismatch; G.gap_open = ctx->ssw_open; G.gap_extend = ctx->ssw_extend;
        G.he = d_he.as<uint32_t>(); G.qbuf = d_qbuf.as<uint8_t>(); G.qcap = maxq;
        G.dir = d_dir.as<uint8_t>(); G.dir_cap = dir_cap;
        G.results = ctx->d_results.as<lsw_result>();
        G.cigar = want_cigar ? ctx->d_cigar.as<uint32_t>() : nullptr; G.cigar_cap = want_cigar ? cigar_cap : 0;
        G.ncigar = ctx->d_ncigar.as<unsigned long long>();
        G.errflag = ctx->d_err.as<int32_t>();
        k_ssw_generic<<<nslots / 128, 128, 0, ctx->stream>>>(G, nslots);
        CKB(cudaGetLastError());
        CKB(cudaMemcpyAsync(out, ctx->d_results.p, (size_t)n * sizeof(lsw_result), cudaMemcpyDeviceToHost, ctx->stream));
        CKB(cudaMemcpyAsync(&h_ncig, ctx->d_ncigar.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CKB(cudaMemcpyAsync(&h_err, ctx->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CKB(cudaStreamSynchronize(ctx->stream));
        if (h_err) { rc = fail(ctx, LSW_E_UNSUPPORTED, "lsw_ssw_subreads: an alignment exceeded the band / path limits of the stage-2 kernel"); break; }
        if (cigar_used) *cigar_used = (int64_t)h_ncig;
        if (want_cigar) {
            if ((int64_t)h_ncig > cigar_cap) { rc = fail(ctx, LSW_E_OVERFLOW, "cigar buffer too small; *cigar_used runs are needed"); break; }
            if (h_ncig) {
                CKB(cudaMemcpyAsync(cigar_buf, ctx->d_cigar.p, (size_t)h_ncig * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CKB(cudaStreamSynchronize(ctx->stream));
            }
        }
#undef CKB
    } while (0);
    cudaStreamSynchronize(ctx->stream);
    release();
    return rc;
}

int lsw_chain_subreads(lsw_ctx* ctx, int n_fwd, const int32_t* fwd_order, int n_rev, const int32_t* rev_order, int t_query,
                       int tc_query, lsw_subread* out, int64_t cap, int64_t* n_out) {
    if (!ctx) return LSW_E_ARG;
    if (n_out) *n_out = 0;
    if (n_fwd < 1 || n_fwd > LSW_MAX_ORDER || n_rev < 1 || n_rev > LSW_MAX_ORDER || !fwd_order || !rev_order || cap < 0 || (cap > 0 && !out))
        return fail(ctx, LSW_E_ARG, "lsw_chain_subreads: bad argument (1..16 names per primer order)");
    ChainOrders C;
    memset(&C, 0, sizeof C);
    C.n_fwd = n_fwd; C.n_rev = n_rev; C.t_query = t_query; C.tc_query = tc_query; C.fwd_t = -1; C.rev_tc = -1;
    for (int i = 0; i < n_fwd; i++) { C.fwd[i] = fwd_order[i]; if (fwd_order[i] == t_query && C.fwd_t < 0) C.fwd_t = i; }
    for (int i = 0; i < n_rev; i++) { C.rev[i] = rev_order[i]; if (rev_order[i] == tc_query && C.rev_tc < 0) C.rev_tc = i; }
    if (t_query < 0 || tc_query < 0 || C.fwd_t < 0 || C.rev_tc < 0)
        return fail(ctx, LSW_E_ARG, "lsw_chain_subreads: 'T' must be in fwd_primer_order and 'Tc' in rev_primer_order (the reference raises ValueError)");
    if (ctx->n_hits == 0 || ctx->n_reads == 0) return LSW_OK;
    if (!ctx->hits_pruned) return fail(ctx, LSW_E_STATE, "lsw_chain_subreads needs the hits of an lsw_find_all that ran with lsw_set_overlap >= 0");
    CK(cudaSetDevice(ctx->device));
    const int64_t n = ctx->n_reads;
    CK(ctx->d_nseg.ensure((size_t)(n + 1) * 4)); CK(ctx->d_segpos.ensure((size_t)(n + 1) * 4));
    CK(cudaMemsetAsync(ctx->d_nseg.p, 0, (size_t)(n + 1) * 4, ctx->stream));
    const int grid = grid_for(n, 128, ctx->num_sms * 8);
    k_chain_subreads<<<grid, 128, 0, ctx->stream>>>(ctx->d_hits_sorted.as<lsw_hit>(), ctx->n_hits, n, ctx->read_base, ctx->d_rlen.as<int32_t>(), C,
                                                    ctx->d_nseg.as<int32_t>(), nullptr, nullptr);
    CK(cudaGetLastError());
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_nseg.as<int32_t>(), ctx->d_segpos.as<int32_t>(), (int)(n + 1), ctx->stream));
    CK(ctx->d_cubtemp.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->d_cubtemp.p, tmp, ctx->d_nseg.as<int32_t>(), ctx->d_segpos.as<int32_t>(), (int)(n + 1), ctx->stream));
    int32_t total = 0;
    CK(cudaMemcpyAsync(&total, ctx->d_segpos.as<int32_t>() + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = total;
    if (total > cap) return fail(ctx, LSW_E_OVERFLOW, "lsw_chain_subreads: caller buffer too small; *n_out records are needed");
    if (total == 0) return LSW_OK;
    CK(ctx->d_res_s.ensure((size_t)total * sizeof(lsw_subread)));
    k_chain_subreads<<<grid, 128, 0, ctx->stream>>>(ctx->d_hits_sorted.as<lsw_hit>(), ctx->n_hits, n, ctx->read_base, ctx->d_rlen.as<int32_t>(), C,
                                                    nullptr, ctx->d_segpos.as<int32_t>(), reinterpret_cast<lsw_subread*>(ctx->d_res_s.p));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_res_s.p, (size_t)total * sizeof(lsw_subread), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LSW_OK;
}

int lsw_fetch_hits(lsw_ctx* ctx, lsw_hit* out, int64_t cap, int64_t* n_out) {
    if (!ctx) return LSW_E_ARG;
    if (n_out) *n_out = ctx->n_hits;
    if (ctx->n_hits > cap) return fail(ctx, LSW_E_OVERFLOW, "lsw_fetch_hits: caller buffer too small");
    if (ctx->n_hits == 0) return LSW_OK;
    if (!out) return fail(ctx, LSW_E_ARG, "lsw_fetch_hits: out is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, ctx->d_hits_sorted.p, (size_t)ctx->n_hits * sizeof(lsw_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LSW_OK;
}

int lsw_fetch_hits_compact(lsw_ctx* ctx, lsw_hit16* out, int64_t cap, int64_t* n_out) {
    if (!ctx) return LSW_E_ARG;
    if (n_out) *n_out = ctx->n_hits;
    if (ctx->n_hits > cap) return fail(ctx, LSW_E_OVERFLOW, "lsw_fetch_hits_compact: caller buffer too small");
    if (ctx->n_hits == 0) return LSW_OK;
    if (!out) return fail(ctx, LSW_E_ARG, "lsw_fetch_hits_compact: out is NULL");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_hits16.ensure((size_t)ctx->n_hits * sizeof(lsw_hit16)));
    k_pack_hits16<<<grid_for(ctx->n_hits, 256, ctx->num_sms * 8), 256, 0, ctx->stream>>>(
        ctx->d_hits_sorted.as<lsw_hit>(), ctx->n_hits, ctx->d_hits16.as<uint4>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_hits16.p, (size_t)ctx->n_hits * sizeof(lsw_hit16), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LSW_OK;
}

int lsw_host_alloc(int64_t bytes, void** out) {
    if (!out || bytes < 0) return LSW_E_ARG;
    *out = nullptr;
    if (cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        g_create_error = "lsw_host_alloc: cudaHostAlloc failed";
        *out = nullptr;
        return LSW_E_CUDA;
    }
    return LSW_OK;
}

void lsw_host_free(void* p) { if (p) cudaFreeHost(p); }

int lsw_align_tasks(lsw_ctx* ctx, int64_t n, const lsw_task* tasks, lsw_result* out,
                    uint32_t* cigar_buf, int64_t cigar_cap, int64_t* cigar_used) {
    if (!ctx) return LSW_E_ARG;
    if (n < 0 || (n > 0 && (!tasks || !out))) return fail(ctx, LSW_E_ARG, "lsw_align_tasks: bad argument");
    if (ctx->nseq <= 0) return fail(ctx, LSW_E_STATE, "lsw_set_queries has not been called");
    if (int rc = check_scoring_fits(ctx)) return rc;
    if (n >= (1ll << 30)) return fail(ctx, LSW_E_ARG, "at most 2^30 - 1 tasks per call");
    if (cigar_used) *cigar_used = 0;
    if (n == 0) return LSW_OK;
    CK(cudaSetDevice(ctx->device));
    read_env(ctx);
    ctx->timed.clear(); ctx->ev_used = 0;

    // group by sequence (K1 serves one sequence per warp): counting sort on the host
    std::vector<int64_t> h_seg((size_t)ctx->nseq + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        const lsw_task& t = tasks[i];
        if (t.query < 0 || t.query >= ctx->nseq || t.read < 0 || t.read >= ctx->n_reads || t.start < 0 ||
            t.end < t.start || t.end > ctx->h_rlen[t.read])
            return fail(ctx, LSW_E_ARG, "lsw_align_tasks: task out of range");
        h_seg[t.query + 1]++;
    }
    for (int s = 0; s < ctx->nseq; s++) h_seg[s + 1] += h_seg[s];
    std::vector<int32_t> hr(n), ha(n), hb(n), ho(n);
    std::vector<int16_t> hs(n);
    {
        std::vector<int64_t> fill(h_seg.begin(), h_seg.end() - 1);
        for (int64_t i = 0; i < n; i++) {
            const lsw_task& t = tasks[i];
            const int64_t d = fill[t.query]++;
            hr[d] = t.read; ha[d] = t.start; hb[d] = t.end; hs[d] = (int16_t)t.query; ho[d] = (int32_t)i;
        }
    }
    TaskBuf& cur = ctx->tb[0];
    CK(cur.ensure((size_t)n, true));
    CK(cudaMemcpyAsync(cur.read.p, hr.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(cur.a.p, ha.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(cur.b.p, hb.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(cur.seq.p, hs.data(), (size_t)n * 2, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(cur.orig.p, ho.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_seg[0].ensure(h_seg.size() * 8));
    CK(cudaMemcpyAsync(ctx->d_seg[0].p, h_seg.data(), h_seg.size() * 8, cudaMemcpyHostToDevice, ctx->stream));

    std::vector<uint8_t> tab;
    if (ctx->ssw) tab.assign((size_t)ctx->nseq * 128, 1);
    else
    build_cand_ok(ctx->nseq, ctx->qlen.data(), ctx->match, ctx->mismatch, ctx->g, 0.0, true, tab);          // every task gets a traceback
    CK(ctx->d_cand_ok.ensure(tab.size()));
    CK(cudaMemcpyAsync(ctx->d_cand_ok.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_counters.ensure(32)); CK(ctx->d_err.ensure(4)); CK(ctx->d_ncigar.ensure(8));
    CK(cudaMemsetAsync(ctx->d_counters.p, 0, 32, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_err.p, 0, 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ncigar.p, 0, 8, ctx->stream));
    CK(ctx->d_results.ensure((size_t)n * sizeof(lsw_result)));
    const bool want_cigar = cigar_buf != nullptr && cigar_cap > 0;
    if (want_cigar) CK(ctx->d_cigar.ensure((size_t)cigar_cap * 4));

    TraceParams T;
    fill_trace_params(ctx, T, cur);
    T.mode = 1;
    T.results = ctx->d_results.p;
    T.cigar = want_cigar ? ctx->d_cigar.as<uint32_t>() : nullptr;
    T.cigar_cap = want_cigar ? cigar_cap : 0;
    T.ncigar = ctx->d_ncigar.as<unsigned long long>();
    int rc;
    if (ctx->ssw) {
        SswParams S;
        memset(&S, 0, sizeof S);
        S.mode = 1; S.results = T.results; S.cigar = T.cigar; S.cigar_cap = T.cigar_cap; S.ncigar = T.ncigar;
        rc = run_round_ssw(ctx, cur, 0, h_seg, n, S, nullptr);
    } else {
        rc = run_round(ctx, cur, 0, h_seg, n, T, nullptr);
    }
    if (rc) return rc;

    unsigned long long h_ncig = 0;
    int32_t h_err = 0;
    CK(cudaMemcpyAsync(out, ctx->d_results.p, (size_t)n * sizeof(lsw_result), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&h_ncig, ctx->d_ncigar.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&h_err, ctx->d_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->timed.clear(); ctx->ev_used = 0;
    if (h_err) return fail(ctx, LSW_E_INTERNAL, "traceback window bound violated");
    if (cigar_used) *cigar_used = (int64_t)h_ncig;
    if (want_cigar) {
        if ((int64_t)h_ncig > cigar_cap) return fail(ctx, LSW_E_OVERFLOW, "cigar buffer too small; *cigar_used runs are needed");
        if (h_ncig) {
            CK(cudaMemcpyAsync(cigar_buf, ctx->d_cigar.p, (size_t)h_ncig * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    return LSW_OK;
}

int lsw_measure_int_peak(lsw_ctx* ctx, int mode, int iters, double* lane_ops_per_s, float* ms_out) {
    if (!ctx || !lane_ops_per_s || iters <= 0) return LSW_E_ARG;
    CK(cudaSetDevice(ctx->device));
    int threads = 256, blocks = ctx->num_sms * 8;
    if (mode >= 9) { threads = SCORE_WARPS * 32; blocks = ctx->num_sms * ScoreOcc<24>::value; }
    const size_t smem9 = sizeof(U2) * (size_t)NIDX * ScoreLane<24>::PROF_STRIDE;
    DevBuf out, in;
    CK(out.ensure((size_t)threads * blocks * 4));
    CK(in.ensure(64));
    const uint32_t h_in[8] = {0x00030002u, 0x00010001u, 0x01000100u, 0x00100020u, 0x00000000u, 1u, 0, 0};
    CK(cudaMemcpyAsync(in.p, h_in, sizeof h_in, cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0, ctx->stream));
#define LSW_IP(M) case M: k_intpeak<M><<<blocks, threads, 0, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters, 1u); break;
        switch (mode) {
            LSW_IP(0) LSW_IP(1) LSW_IP(2) LSW_IP(3) LSW_IP(4) LSW_IP(5) LSW_IP(6) LSW_IP(7) LSW_IP(8)
            case 9: k_intpeak_k1row<LSW_K1_KEYV><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;   // what K1 is built with
            case 10: k_intpeak_k1row<0><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;
            case 11: k_intpeak_k1row<1><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;
            case 12: k_intpeak_k1row<2><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;
            case 13: k_intpeak_k1row<3><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;
            case 14: k_intpeak_k1row<4><<<blocks, threads, smem9, ctx->stream>>>(out.as<uint32_t>(), in.as<uint32_t>(), iters); break;
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); out.release(); in.release();
                     return fail(ctx, LSW_E_ARG, "lsw_measure_int_peak: unknown mode");
        }
#undef LSW_IP
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out.release(); in.release();
    // mode 9: one iteration = column2 = 24 rows x (5 DPX + 4 adds) lane-ops = 96 cells per lane
    const double ops = (double)threads * blocks * (double)iters * (mode >= 9 ? 24 * 9 : intpeak_ops_per_iter(mode));
    *lane_ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return LSW_OK;
}

}  // extern "C"
