// lsw_common.cuh -- shared definitions for the B200 Smith-Waterman seeding kernels.
//
// Everything here is written so that the per-lane logic can ALSO be compiled for the host
// (nvcc host pass) and driven one lane at a time by csrc/lsw_emu.cu.  That emulation is test
// infrastructure (tests/test_emu.py checks the kernel logic against the oracle without a GPU);
// the product library (lsw.cu) never runs it.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define LSW_HD __host__ __device__ __forceinline__

namespace lsw {

constexpr int SC_SHIFT   = 8;          // DP values are scaled by 256 inside an s16 half ...
constexpr int SC         = 1 << SC_SHIFT;   // ... so the low byte can carry the step index of the arg-max key
constexpr int BLOCK_STEPS = 64;        // columns a lane advances between two task-switch points
constexpr int MAXQ       = 63;         // LSW_MAX_QUERY_LEN
constexpr int BB_BLOCK   = 32;         // columns per stored block best of round 0 (lsw_reuse.cuh)
constexpr int BB_GROUP   = 32;         // block bests per second-level entry (long slices are merged group by group)
constexpr int QSTRIDE    = 64;         // bytes per query in the device query-code table
constexpr int CODE_OTHER = 4;          // read code of any byte that is not A/C/G/T (never matches)
constexpr int CODE_OTHER_HI = 20;      // ... as seen by the task in the HIGH half of a K1 lane (see profile index)
// K1 profile row of the pair (code of the low-half task, code of the high-half task) is
//   idx = 4 * lo + hi            with lo in {0..3, 4 = other} and hi in {0..3, 20 = other}:
//   0..15 ACGT x ACGT, 16..19 lo other, 20/24/28/32 hi other, 36 both other.  The sixteen common
//   rows are consecutive so that, with a row stride of an odd number of 8-byte words, a half-warp's
//   LDS.64 requests fall into sixteen distinct bank pairs (no shared-memory bank conflicts).
constexpr int NIDX       = 37;

// ---- packed s16x2 DPX primitives (VIADDMNMX.S16x2 / VIMNMX3.S16x2 on sm_100a) ------------
LSW_HD uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) {       // max(a + b, c) per s16 half
    return __viaddmax_s16x2(a, b, c);
}
LSW_HD uint32_t max3_2(uint32_t a, uint32_t b, uint32_t c) {        // max(a, b, c) per s16 half
    return __vimax3_s16x2(a, b, c);
}
LSW_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)((v >> (8 * (s & 7))) & 0xFF);
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// swalign upper-cases both strings and compares raw characters: A/C/G/T in
// either case get codes 0..3, every other byte gets CODE_OTHER and never matches an ACGT query.
LSW_HD uint8_t encode_base(uint8_t ch) {
    if (ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 32);
    return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : (uint8_t)CODE_OTHER;
}

LSW_HD unsigned long long atomic_add_ull(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const unsigned long long o = *p; *p += v; return o;    // host emulation is single-threaded
#endif
}
// counter++ for every calling lane with ONE atomic per warp: the lanes that are converged here elect a
// leader, which adds their count; each gets its own slot.  (Hit and candidate appends all target a single
// counter; per-lane atomics serialise on that address in L2.)
LSW_HD unsigned long long append_slot(unsigned long long* counter) {
#if defined(__CUDA_ARCH__)
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
#else
    return (*counter)++;
#endif
}

// same, for lanes that may target DIFFERENT counters (one per class): lanes with the same counter are grouped
LSW_HD unsigned long long append_slot_grouped(unsigned long long* counter) {
#if defined(__CUDA_ARCH__)
    const unsigned mask = __match_any_sync(__activemask(), (unsigned long long)counter);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
#else
    return (*counter)++;
#endif
}

LSW_HD void atomicExch_or_store(int32_t* p, int32_t v) { *p = v; }   // benign race: any writer stores the same flag

// ---- task / result records shared by host and device -------------------------------------
struct TaskArrays {           // struct-of-arrays task list of one recursion round
    int32_t* read;            // read index in the uploaded batch
    int32_t* a;               // slice start on the read
    int32_t* b;               // slice end (exclusive)
    int16_t* seq;             // schema sequence index
    int32_t* orig;            // align_tasks(): caller's task index (results are written there); else unused
    uint8_t* skip;            // K1 segments (lsw_reuse.cuh): leading 64-column blocks that only warm the DP up; else null
};

// What K1 needs to start a task, in one aligned 16-byte load (built when the task list is built, so a lane
// does not chase read -> offset -> codes through three dependent loads at every task switch).
// What K2 needs to trace a candidate, in one 32-byte record written where the candidate is queued (K1's finish /
// the merge kernel): K2 visits candidates in (sequence, window length) order, i.e. in random task order, and
// would otherwise gather seven scattered fields per task.
struct __align__(16) CandRec {
    int64_t p0;               // offset of the slice start in the code array
    int32_t len;              // slice length
    int32_t rd, a;            // read index and slice start (hits and children are reported in read coordinates)
    int32_t res_x, col;       // score | max_row << 16, max_col (slice relative)
    int32_t t;                // task index (child slots, fallback list, caller's result slot)
};

struct __align__(16) KTask {
    int64_t p0;               // byte offset in codes of the slice's first base
    int32_t len;              // slice length
    int32_t skip;             // leading lane blocks that only warm the DP up (tail segments of lsw_reuse.cuh)
};

struct ScoreParams {
    const uint8_t*  codes;    // read codes, one byte per base, every read 16-byte aligned
    const int64_t*  roff;     // [n_reads + 1] byte offset of each read in codes
    TaskArrays      t;
    const KTask*    kt;       // [n_tasks] start descriptors, parallel to t
    const int64_t*  seg;      // [n_seq + 1] tasks of sequence s are [seg[s], seg[s+1])
    unsigned long long* head; // [n_seq] next unclaimed task of each sequence (relative to seg[s])
    int2*           res;      // [n_tasks] x = score | (max_row << 16), y = max_col (1-based DP coords)
    const uint8_t*  qcodes;   // [n_seq][QSTRIDE] 0..3
    const int32_t*  qlen;     // [n_seq]
    const int32_t*  class_seqs; // sequences served by this launch (same register-row class)
    int             n_class_seqs;
    const uint8_t*  cand_ok;  // [n_seq][128] non-zero if a task with that score can still pass the hit test (lsw_host.hpp)
    uint32_t*       cand;     // out: tasks that need a traceback
    uint16_t*       cand_key; // out: sequence << 8 | K2 window length (sort key; preset to 0xFFFF)
    CandRec*        cand_rec; // out: the candidates themselves; cand[] holds indices into this array
    unsigned long long* ncand;
    unsigned long long* counters;   // [0] += reference tasks, [1] += reference cells, [2] += lane steps, [3] += computed cells
    int32_t s_match;          // SC * (match + |gap|)
    int32_t s_mis;            // SC * (mismatch + |gap|)
    int32_t g;                // |gap|
    int32_t match;
    uint32_t* blockbest;      // round 0: best cell of every 64-column block, [n_seq][bb_total] (lsw_reuse.cuh); else null
    const int64_t* bb_off;    // [n_reads] first block of each read
    int64_t bb_total;
    int32_t push_cands;       // 1: finish() queues traceback candidates; 0: the merge kernel does (segment mode)
    int32_t count_ref;        // 1: the tasks are reference align() calls (count them); 0: segments
    int32_t block_steps;      // columns between two task-switch points: BLOCK_STEPS, or 32 for segment lists
    int32_t aff_open, aff_extend;   // lsw_affine.cuh (skbio path): gap open / gap extension penalties, positive; s_match / s_mis are 256 * score there
};

struct TraceParams {
    const uint8_t*  codes;
    const int64_t*  roff;
    TaskArrays      t;
    const int2*     res;
    const uint8_t*  qcodes;
    const int32_t*  qlen;
    const uint8_t*  cand_ok;      // [n_seq][128] 0, or 1 + the most errors a passing alignment with that score can have (lsw_host.hpp)
    const uint32_t* cand;         // candidate task ids, sorted by (sequence, window length)
    const uint16_t* cand_key;     // their sort keys: sequence << 8 | window length (0xFFFF = padding)
    const CandRec*  cand_rec;     // candidate records, indexed by cand[]
    const unsigned long long* ncand;
    unsigned long long* head;     // [n_seq] next unclaimed candidate of each sequence
    const int32_t*  class_seqs;   // sequences of this launch's register-row class
    int             n_class_seqs;
    // tiers: k_trace_band (shared-memory band) -> k_trace_fast (whole window in a global scratch) -> k_trace_full
    uint32_t*       fb2;          // out of the band kernel: candidate records that left the band, per sequence at the sequence's offset in cand[]
    unsigned long long* fb2_cnt;  // [n_seq] their number
    const uint32_t* fb_in;        // in of k_trace_fast: null = every candidate (cand[]), else the band kernel's fb2 / fb2_cnt
    const unsigned long long* fb_in_cnt;
    uint32_t*       fallback;     // tasks the short-window kernels could not prove exact
    unsigned long long* nfallback;
    uint32_t*       scratch;      // fast: H bytes, [block][window col][unit][thread] x 16 B; full: op bits, [window col][word][thread]
    int64_t         scratch_stride;   // words between consecutive (column, word) rows: threads of a block (fast) / of the launch (full)
    int64_t         scratch_block_words;  // k_trace_fast: 16-byte scratch elements per block
    int32_t match, mismatch, g;
    int32_t mode;             // 0 = find_all (hit test, children), 1 = align_tasks (result records, CIGAR)
    double  thr;
    int32_t round;
    // mode 0 outputs
    void*   hits;             // lsw_hit[]
    int64_t hits_cap;
    unsigned long long* nhits;
    uint8_t*  child_flag;     // [2 * n_tasks]
    TaskArrays child;         // [2 * n_tasks]
    // mode 1 outputs
    void*   results;          // lsw_result[] indexed by t.orig
    uint32_t* cigar;
    int64_t cigar_cap;
    unsigned long long* ncigar;
    int32_t* errflag;
    int32_t all_fallback;     // 1: the scoring is outside the short-window kernel's range (nibble_walk_ok), every candidate takes the full window
    int32_t debug;            // timing experiments only (LSW_K2_DEBUG): 1 = fill only, 2 = fill without the scratch stores, 4 = no output
};

}  // namespace lsw
