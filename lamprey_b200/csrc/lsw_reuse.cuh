// lsw_reuse.cuh -- reuse of the round-0 DP by the recursion's child tasks.
//
// findPrimerAlignments (/root/reference/lamprey.py:596-645) re-aligns the primer against what lies left and
// right of every hit, so most cells of a child slice [a, b) were already computed when the whole read was
// aligned in round 0.  They are the SAME cells: an alignment whose every prefix scores > 0 spans fewer than
// Wx = Q + M*Q/g read columns, so for read positions p >= a + Wx the child's H(r, p) equals the whole read's
// H(r, p) (both have a zero boundary at their start, and no alignment ending at p can reach back to it).
// Round 0 therefore stores, for every BB_BLOCK = 32 columns of every (read, sequence), the block's best cell
// (value, row, column) -- 4 bytes -- and a child task whose slice is long enough is computed as
//     HEAD   [a, hb)            the child's own DP up to the first block boundary hb >= a + Wx
//     MIDDLE blocks hb/32 .. te/32-1 of the whole read, looked up                  (te = last boundary <= b)
//     TAIL   [te - Wxa, b)      a DP that warms up for Wxa >= Wx columns (not counted) and counts (te, b];
//                               Wxa is a multiple of the 32-column lane block the segment launches use
// and the three bests are merged with swalign's tie rule (value, then row, then the LAST column).
// In lsw_find_all only cells that could pass the hit test have to be the same cells: Wx shrinks to Q + (M*Q - minpass)/g + 1
// (lsw_host.hpp build_reuse_wx, with the argument), about 1.9 Q instead of 3 Q for LAMPrey's scoring and thresholds.
// Everything here is exact for every hit and every child; nothing is approximated.  Slices too short to profit are computed directly.
#pragma once
#include "lsw_common.cuh"

namespace lsw {

constexpr int SEG_BLOCK = 16;             // lane block (columns between task-switch points) of segment launches
constexpr int REUSE_MIN_SAVING = 16;      // columns a split must save (measured: 96 -> 16 takes 1.5 % off K1; below that nothing changes)

struct SegPlan {
    int nseg;        // K1 segments of the task: 1 (direct, or head only) or 2 (head + tail)
    bool split;
    int hb;          // head = [a, hb)
    int te;          // last block boundary <= b
    int mid_end;     // middle = whole-read blocks hb/BB_BLOCK .. mid_end - 1
    int ts, skip;    // tail = [ts, b), its first `skip` lane blocks (SEG_BLOCK columns each) are warm-up
    bool tail;
};

// bbrow: the (read, sequence) row of the round-0 block bests (may be null: no look-ahead into the last block).
// The last, partial block [te, b) needs no DP when the whole block's best cell lies before b -- it is then also the best
// of [te, b), last-column tie rule included -- or when the block holds no positive cell at all.
LSW_HD SegPlan plan_task(int a, int b, int Q, int match, int g, bool enable, const uint32_t* bbrow = nullptr,
                         int tb = SEG_BLOCK /* lane block of the segment launch */, int wx = 0 /* head length (build_reuse_wx); 0: a priori */) {
    SegPlan p;
    p.nseg = 1; p.split = false; p.hb = b; p.te = b; p.mid_end = 0; p.ts = 0; p.skip = 0; p.tail = false;
    if (!enable) return p;
    const int Wx = wx > 0 ? wx : Q + (match * Q) / g;
    const int hb = ((a + Wx + BB_BLOCK - 1) / BB_BLOCK) * BB_BLOCK;
    const int te = (b / BB_BLOCK) * BB_BLOCK;
    if (hb > te) return p;
    const int Wxa = ((Wx + tb - 1) / tb) * tb;
    int ts = te - Wxa;
    if (ts < 0) ts = 0;
    bool tail = b > te;
    int mid_end = te / BB_BLOCK;
    if (tail && bbrow) {
        const uint32_t v = bbrow[te / BB_BLOCK];
        const int s = (int)(v >> 16) - g;
        const int col1 = BLOCK_STEPS * (te / BLOCK_STEPS) + (int)(v & 63) + 1;      // 1-based read column of the block's best cell
        if (s <= 0 || col1 <= b) { tail = false; mid_end++; }
    }
    const int computed = (hb - a) + (tail ? b - ts : 0);
    if (computed + REUSE_MIN_SAVING > b - a) return p;
    p.split = true; p.hb = hb; p.te = te; p.mid_end = mid_end; p.ts = ts; p.skip = (te - ts) / tb; p.tail = tail;
    p.nseg = tail ? 2 : 1;
    return p;
}

struct ReuseParams {
    TaskArrays t;                // child tasks of the round
    int64_t n;
    const int32_t* qlen;
    int32_t match, g, enable;
    int32_t* nseg;               // plan: segments per task (scanned into segpos)
    const int32_t* segpos;       // [n + 1]
    KTask* s_kt;                 // its start descriptors
    const int64_t* roff;         // [n_reads] byte offset of each read in codes
    const int2* res_s;           // K1 results per segment
    int2* res_t;                 // merged result per task
    const uint32_t* blockbest;   // [n_seq][bb_total]
    const int64_t* bb_off;       // [n_reads + 1] first block of each read
    int64_t bb_total;
    const uint32_t* bb1;         // second level, [n_seq][bb1_total]: best of every BB_GROUP blocks of a read (null: not built)
    const int64_t* bb1_off;      // [n_reads + 1] first group of each read
    int64_t bb1_total;
    const uint8_t* cand_ok;      // [n_seq][128]
    const int32_t* wx;           // [n_seq] head / warm-up length of a child slice (lsw_host.hpp build_reuse_wx); null: the a-priori bound
    const int32_t* class_of;     // [n_seq] register-row class
    uint32_t* cand[16];          // per class: candidate list / sort key
    uint16_t* cand_key[16];
    CandRec* cand_rec[16];
    unsigned long long* ncand;   // [classes]
    unsigned long long* counters;   // [0] += tasks, [1] += reference cells
};

LSW_HD SegPlan plan_of(const ReuseParams& R, int64_t t) {
    const int q = R.t.seq[t];
    return plan_task(R.t.a[t], R.t.b[t], R.qlen[q], R.match, R.g, R.enable != 0,
                     R.blockbest ? R.blockbest + (size_t)q * R.bb_total + R.bb_off[R.t.read[t]] : nullptr, SEG_BLOCK,
                     R.wx ? R.wx[q] : 0);
}

// K1 needs nothing but the start descriptor of a segment (its results go to res_s[segment], the task list is not read
// in segment mode), so that is all that is written.
LSW_HD void fill_segments(const ReuseParams& R, int64_t t) {
    const int a = R.t.a[t], b = R.t.b[t];
    const SegPlan p = plan_of(R, t);
    const int64_t i = R.segpos[t];
    const int64_t r0 = R.roff[R.t.read[t]];
    KTask k; k.p0 = r0 + a; k.len = (p.split ? p.hb : b) - a; k.skip = 0;
    R.s_kt[i] = k;
    if (p.split && p.tail) {
        k.p0 = r0 + p.ts; k.len = b - p.ts; k.skip = p.skip;
        R.s_kt[i + 1] = k;
    }
}

// A stored block best -> (score, row, 1-based read column).  j: block index within the read.
LSW_HD void decode_block_best(uint32_t v, int j, int g, int& s, int& r, int& c1) {
    s = (int)(v >> 16) - g;                            // the column is stored as the step within K1's 64-step lane block
    r = (int)((v >> 8) & 0xFF) + (int)((v >> 6) & 1) + 1;
    c1 = BLOCK_STEPS * (j * BB_BLOCK / BLOCK_STEPS) + (int)(v & 63) + 1;
}

// Second level: the best of blocks [BB_GROUP * G, BB_GROUP * (G + 1)) of one (read, sequence) under swalign's tie rule
// (value, then row, then the LAST column -- the blocks are visited in ascending order and >= keeps the later one), as
// score << 16 | row << 10 | column offset within the group's BB_GROUP * BB_BLOCK = 1024 columns.  0: no positive cell.
LSW_HD uint32_t reduce_block_group(const uint32_t* bbrow, int nblocks, int G, int g) {
    int score = 0, row = 0, col1 = 0;
    const int j1 = (G + 1) * BB_GROUP < nblocks ? (G + 1) * BB_GROUP : nblocks;
    for (int j = G * BB_GROUP; j < j1; j++) {
        int s, r, c1;
        decode_block_best(bbrow[j], j, g, s, r, c1);
        if (s > 0 && (s > score || (s == score && r >= row))) { score = s; row = r; col1 = c1; }
    }
    if (score <= 0) return 0u;
    return ((uint32_t)score << 16) | ((uint32_t)row << 10) | (uint32_t)(col1 - 1 - G * BB_GROUP * BB_BLOCK);
}

// sort key of a K2 candidate: the number of window columns its traceback will fill
LSW_HD int cand_window(int score, int row, int col, int match, int g) {
    int w = score > 0 ? row + (match * row - score) / g + 8 : 0;
    return w > col ? col : w;
}

// merged (score, max_row, max_col) of task t; returns true if it must be traced (candidate)
LSW_HD bool combine_task(const ReuseParams& R, int64_t t, int& o_score, int& o_row, int& o_col) {
    const int a = R.t.a[t], b = R.t.b[t], q = R.t.seq[t], rd = R.t.read[t];
    const int Q = R.qlen[q];
    const SegPlan p = plan_of(R, t);
    const int64_t i = R.segpos[t];
    int score = 0, row = 0, col = 0;
    auto consider = [&](int s, int r, int c) {       // regions are visited in ascending column order: >= keeps the last
        if (s > 0 && (s > score || (s == score && r >= row))) { score = s; row = r; col = c; }
    };
    {
        const int2 h = R.res_s[i];
        consider(h.x & 0xFFFF, h.x >> 16, h.y);
    }
    if (p.split) {
        const uint32_t* bb = R.blockbest + (size_t)q * R.bb_total + R.bb_off[rd];
        auto block = [&](int j) {
            int s, r, c1;
            decode_block_best(bb[j], j, R.g, s, r, c1);
            consider(s, r, c1 - a);
        };
        int j = p.hb / BB_BLOCK;
        const int j1 = p.mid_end;
        if (R.bb1) {
            // whole groups in the middle of a long slice come from the second level
            const uint32_t* b1 = R.bb1 + (size_t)q * R.bb1_total + R.bb1_off[rd];
            int ja = ((j + BB_GROUP - 1) / BB_GROUP) * BB_GROUP;
            if (ja > j1) ja = j1;
            for (; j < ja; j++) block(j);
            for (; j + BB_GROUP <= j1; j += BB_GROUP) {
                const int G = j / BB_GROUP;
                const uint32_t v = b1[G];
                consider((int)(v >> 16), (int)((v >> 10) & 63), G * BB_GROUP * BB_BLOCK + (int)(v & 1023) + 1 - a);
            }
        }
        for (; j < j1; j++) block(j);
        if (p.tail) {
            const int2 h = R.res_s[i + 1];
            consider(h.x & 0xFFFF, h.x >> 16, p.ts + h.y - a);
        }
    }
    if (score == 0) { row = b > a ? Q : 0; col = b > a ? b - a : 0; }
    R.res_t[t] = make_int2(score | (row << 16), col);
    o_score = score; o_row = row; o_col = col;
    return R.cand_ok[(size_t)q * 128 + score] != 0;
}

}  // namespace lsw
