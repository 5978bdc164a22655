// lsw_pool.cpp -- several contexts, one host thread each, behind one handle (include/lsw.h, "Pool").
//
// Host-side only, written against the public C ABI of lsw.cu: a worker thread owns one lsw_ctx and runs
// lsw_upload_reads -> lsw_find_all -> lsw_fetch_hits[_compact] for the read range it is handed.  Mirrors what
// /root/reference/lamprey.py:2182-2207 does with processes: deal the reads over workers, collect the results in order.
//   * one batch is cut into contiguous read ranges balanced by BASES, one per device (an ONT length distribution has a
//     10x tail: equal read counts would not be equal work);
//   * a range's hits go straight from its device into their final place in the caller's buffer: the offset is the sum of the
//     hit counts of the ranges before it, which every worker waits for after its own lsw_find_all (the "gather");
//   * per device one upload, one lsw_find_all and one download run at a time; with two contexts per device the copies of one
//     batch overlap the kernels of the next.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lsw.h"

namespace {

std::string g_pool_create_error;

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Batch {                       // one lsw_pool_submit
    int64_t ticket = 0;
    const uint8_t* ascii = nullptr;
    const int64_t* offs = nullptr;
    std::vector<int64_t> cut;        // [n_ranges + 1] read ranges (n_ranges = devices x pieces; range r runs on device r % devices)
    uint8_t* out = nullptr;
    int64_t cap = 0;
    int item = 32;
    // filled by the workers (under Pool::mu)
    std::vector<int64_t> nhits;      // per range, -1 = not known yet
    std::vector<int> done;           // per range
    std::vector<lsw_counters> cnt;
    std::vector<double> t_up, t_find, t_fetch, t_wait;
    int rc = LSW_OK;
    std::string err;
    double t_submit = 0, t_done = 0;
    int left = 0;
};

// Per device: uploads and lsw_find_all calls run one at a time AND in submission order (a turnstile each: a job takes its number at
// submit).  Plain mutexes let the second of two batches submitted back to back overtake the first at the very first lock, after
// which the caller -- who waits for tickets in order -- sat on a finished batch while the pipeline ran dry.
struct DeviceLocks {
    std::mutex m;
    std::condition_variable cv;
    int64_t issued = 0, next_h2d = 0, next_gpu = 0;
    std::mutex d2h;
    void enter(int64_t& turn, int64_t seq) { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return turn == seq; }); }
    void leave(int64_t& turn) { { std::lock_guard<std::mutex> g(m); turn++; } cv.notify_all(); }
};

struct Job { std::shared_ptr<Batch> b; int range = 0; int64_t seq = 0; };

}  // namespace

struct lsw_pool {
    std::vector<int> devices;
    int per_dev = 1;
    int pieces = 1;                                        // ranges per device and batch (lsw_pool_set_pieces)
    struct Worker {
        lsw_ctx* ctx = nullptr;
        int dev_index = 0;
        std::thread th;
        std::deque<Job> q;
    };
    std::vector<std::unique_ptr<Worker>> workers;         // [device][slot] flattened: device * per_dev + slot
    std::vector<std::unique_ptr<DeviceLocks>> locks;
    std::mutex mu;                                         // queues, batches, tickets
    std::condition_variable cv_work, cv_done;
    bool stop = false;
    int64_t next_ticket = 1;
    std::map<int64_t, std::shared_ptr<Batch>> live;
    double thr = 0.7;
    int hit_format = LSW_HITS_WIDE;
    bool configured = false;
    std::string err;
};

namespace {

void fail_batch(Batch& b, int rc, const std::string& msg) {
    if (b.rc == LSW_OK) { b.rc = rc; b.err = msg; }
}

void run_job(lsw_pool* p, lsw_pool::Worker* w, const Job& job) {
    Batch& b = *job.b;
    const int r = job.range;
    const int64_t lo = b.cut[r], hi = b.cut[r + 1];
    DeviceLocks& L = *p->locks[w->dev_index];
    int rc = LSW_OK;
    int64_t nh = 0;
    lsw_counters cnt;
    std::memset(&cnt, 0, sizeof cnt);
    double t_up = 0, t_find = 0, t_fetch = 0, t_wait = 0;
    L.enter(L.next_h2d, job.seq);
    if (hi > lo) {
        const double t0 = now_s();
        rc = lsw_set_read_base(w->ctx, lo);
        if (rc == LSW_OK) rc = lsw_upload_reads(w->ctx, hi - lo, b.ascii, b.offs + lo);
        t_up = now_s() - t0;
    }
    L.leave(L.next_h2d);
    L.enter(L.next_gpu, job.seq);
    if (hi > lo && rc == LSW_OK) {
        const double t0 = now_s();
        rc = lsw_find_all(w->ctx, p->thr, &nh, &cnt);
        if (rc == LSW_E_OVERFLOW) {              // the device-side hit buffer was too small: once more with room for the count
            if (lsw_set_hit_capacity(w->ctx, nh + 1024) == LSW_OK) rc = lsw_find_all(w->ctx, p->thr, &nh, &cnt);
            lsw_set_hit_capacity(w->ctx, 0);
        }
        t_find = now_s() - t0;
    }
    L.leave(L.next_gpu);
    int64_t before = 0;
    {
        // the gather: this range's hits start where the ranges before it end
        std::unique_lock<std::mutex> g(p->mu);
        if (rc != LSW_OK) fail_batch(b, rc, std::string("device ") + std::to_string(p->devices[w->dev_index]) + ": " + lsw_last_error(w->ctx));
        b.nhits[r] = rc == LSW_OK ? nh : 0;
        b.cnt[r] = cnt;
        p->cv_done.notify_all();
        const double t0 = now_s();
        p->cv_done.wait(g, [&] { for (int k = 0; k < r; k++) if (b.nhits[k] < 0) return false; return true; });
        t_wait = now_s() - t0;
        for (int k = 0; k < r; k++) before += b.nhits[k];
    }
    if (rc == LSW_OK && nh > 0) {
        if (before + nh > b.cap) rc = LSW_E_OVERFLOW;      // reported by lsw_pool_wait with the needed count
        else {
            std::lock_guard<std::mutex> g(L.d2h);
            const double t0 = now_s();
            int64_t got = 0;
            uint8_t* dst = b.out + (size_t)before * (size_t)b.item;
            rc = p->hit_format == LSW_HITS_COMPACT ? lsw_fetch_hits_compact(w->ctx, reinterpret_cast<lsw_hit16*>(dst), b.cap - before, &got)
                                                   : lsw_fetch_hits(w->ctx, reinterpret_cast<lsw_hit*>(dst), b.cap - before, &got);
            t_fetch = now_s() - t0;
        }
    }
    {
        std::lock_guard<std::mutex> g(p->mu);
        if (rc == LSW_E_OVERFLOW) fail_batch(b, rc, "hit buffer too small for the batch; *n_hits records are needed");
        else if (rc != LSW_OK) fail_batch(b, rc, std::string("device ") + std::to_string(p->devices[w->dev_index]) + ": " + lsw_last_error(w->ctx));
        b.t_up[r] = t_up; b.t_find[r] = t_find; b.t_fetch[r] = t_fetch; b.t_wait[r] = t_wait;
        b.done[r] = 1;
        if (--b.left == 0) b.t_done = now_s();
        p->cv_done.notify_all();
    }
}

void worker_main(lsw_pool* p, lsw_pool::Worker* w) {
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> g(p->mu);
            p->cv_work.wait(g, [&] { return p->stop || !w->q.empty(); });
            if (w->q.empty()) return;                     // stop, and nothing left to do
            job = w->q.front();
            w->q.pop_front();
        }
        run_job(p, w, job);
    }
}

// contiguous read ranges balanced by bases (the product's shard rule: lamprey_b200.seeding.shard_reads computes the same cuts)
std::vector<int64_t> shard_by_bases(const int64_t* offs, int64_t n, int parts) {
    std::vector<int64_t> cut(1, 0);
    const int64_t total = n ? offs[n] - offs[0] : 0;
    for (int r = 1; r < parts; r++) {
        const int64_t target = offs[0] + (total * r) / parts;
        int64_t k = std::lower_bound(offs, offs + n + 1, target) - offs;
        k = std::min<int64_t>(n, std::max<int64_t>(cut.back(), k));
        cut.push_back(k);
    }
    cut.push_back(n);
    return cut;
}

}  // namespace

extern "C" {

const char* lsw_pool_last_error(const lsw_pool* pool) { return pool ? pool->err.c_str() : g_pool_create_error.c_str(); }

int lsw_pool_create(int n_devices, const int* devices, int contexts_per_device, lsw_pool** out) {
    if (!out) { g_pool_create_error = "lsw_pool_create: out is NULL"; return LSW_E_ARG; }
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || !devices || contexts_per_device < 1 || contexts_per_device > 4) {
        g_pool_create_error = "lsw_pool_create: need 1..64 devices and 1..4 contexts per device";
        return LSW_E_ARG;
    }
    std::unique_ptr<lsw_pool> p(new lsw_pool());
    p->devices.assign(devices, devices + n_devices);
    p->per_dev = contexts_per_device;
    for (int d = 0; d < n_devices; d++) {
        p->locks.emplace_back(new DeviceLocks());
        for (int s = 0; s < contexts_per_device; s++) {
            std::unique_ptr<lsw_pool::Worker> w(new lsw_pool::Worker());
            w->dev_index = d;
            const int rc = lsw_create(devices[d], &w->ctx);
            if (rc != LSW_OK) {
                g_pool_create_error = std::string("lsw_pool_create: ") + lsw_last_error(nullptr);
                for (auto& o : p->workers) lsw_destroy(o->ctx);
                return rc;
            }
            p->workers.push_back(std::move(w));
        }
    }
    for (auto& w : p->workers) w->th = std::thread(worker_main, p.get(), w.get());
    *out = p.release();
    return LSW_OK;
}

void lsw_pool_destroy(lsw_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(p->mu);
        p->stop = true;
    }
    p->cv_work.notify_all();
    for (auto& w : p->workers) if (w->th.joinable()) w->th.join();
    for (auto& w : p->workers) lsw_destroy(w->ctx);
    delete p;
}

static int pool_configure(lsw_pool* p, bool ssw, int match, int mismatch, int gap_a, int gap_b, int n_queries, const uint8_t* concat,
                          const int32_t* offs, double thr, int allowed_overlap, int hit_format) {
    if (!p) return LSW_E_ARG;
    std::unique_lock<std::mutex> g(p->mu);
    if (!p->live.empty()) { p->err = "lsw_pool_configure: batches are in flight (wait for every ticket first)"; return LSW_E_STATE; }
    if (hit_format != LSW_HITS_WIDE && hit_format != LSW_HITS_COMPACT) { p->err = "lsw_pool_configure: unknown hit format"; return LSW_E_ARG; }
    if (!(thr > 0.0)) { p->err = "identity_threshold must be > 0 (the reference recurses forever otherwise)"; return LSW_E_ARG; }
    p->configured = false;
    for (auto& w : p->workers) {
        int rc = ssw ? lsw_set_ssw(w->ctx, match, mismatch, gap_a, gap_b) : lsw_set_scoring(w->ctx, match, mismatch, gap_a, gap_b);
        if (rc == LSW_OK) rc = lsw_set_queries(w->ctx, n_queries, concat, offs);
        if (rc == LSW_OK) rc = lsw_set_overlap(w->ctx, allowed_overlap);
        if (rc != LSW_OK) { p->err = lsw_last_error(w->ctx); return rc; }
    }
    p->thr = thr;
    p->hit_format = hit_format;
    p->configured = true;
    return LSW_OK;
}

int lsw_pool_configure(lsw_pool* p, int match, int mismatch, int gap, int gap_ext, int n_queries, const uint8_t* concat,
                       const int32_t* offs, double thr, int allowed_overlap, int hit_format) {
    return pool_configure(p, false, match, mismatch, gap, gap_ext, n_queries, concat, offs, thr, allowed_overlap, hit_format);
}

int lsw_pool_configure_ssw(lsw_pool* p, int match, int mismatch, int gap_open, int gap_extend, int n_queries, const uint8_t* concat,
                           const int32_t* offs, double thr, int allowed_overlap, int hit_format) {
    return pool_configure(p, true, match, mismatch, gap_open, gap_extend, n_queries, concat, offs, thr, allowed_overlap, hit_format);
}

int lsw_pool_set_pieces(lsw_pool* p, int pieces) {
    if (!p) return LSW_E_ARG;
    std::lock_guard<std::mutex> g(p->mu);
    if (pieces < 1 || pieces > 64) { p->err = "lsw_pool_set_pieces: 1..64 pieces per device"; return LSW_E_ARG; }
    p->pieces = pieces;
    return LSW_OK;
}

int lsw_pool_submit(lsw_pool* p, int64_t n, const uint8_t* ascii, const int64_t* offs, void* hits_out, int64_t hits_cap, int64_t* ticket) {
    if (!p) return LSW_E_ARG;
    if (n < 0 || !offs || !ticket || hits_cap < 0 || (hits_cap > 0 && !hits_out) || (n > 0 && !ascii && offs[n] > offs[0])) {
        std::lock_guard<std::mutex> g(p->mu);
        p->err = "lsw_pool_submit: bad argument";
        return LSW_E_ARG;
    }
    std::shared_ptr<Batch> b(new Batch());
    const int D = (int)p->devices.size();
    int S;
    {
        std::lock_guard<std::mutex> g(p->mu);
        S = p->pieces;
    }
    const int R = D * S;
    b->ascii = ascii; b->offs = offs; b->out = static_cast<uint8_t*>(hits_out); b->cap = hits_cap;
    b->cut = shard_by_bases(offs, n, R);
    b->nhits.assign(R, -1); b->done.assign(R, 0); b->cnt.resize(R);
    b->t_up.assign(R, 0); b->t_find.assign(R, 0); b->t_fetch.assign(R, 0); b->t_wait.assign(R, 0);
    b->left = R;
    b->t_submit = now_s();
    {
        std::lock_guard<std::mutex> g(p->mu);
        if (!p->configured) { p->err = "lsw_pool_submit: lsw_pool_configure has not been called"; return LSW_E_STATE; }
        b->item = p->hit_format == LSW_HITS_COMPACT ? (int)sizeof(lsw_hit16) : (int)sizeof(lsw_hit);
        b->ticket = p->next_ticket++;
        p->live[b->ticket] = b;
        // range r runs on device r % D: the first D ranges (one per device) finish first, so the hit counts a later range waits
        // for in the gather are those of earlier waves; on a device, consecutive ranges alternate between its contexts, so that one
        // range's copies overlap the next range's kernels even within a single batch
        for (int r = 0; r < R; r++) {
            Job j; j.b = b; j.range = r;
            const int d = r % D, slot = (int)(((b->ticket - 1) + r / D) % p->per_dev);
            j.seq = p->locks[d]->issued++;           // (under p->mu: the order of submission)
            p->workers[(size_t)d * p->per_dev + slot]->q.push_back(j);
        }
        *ticket = b->ticket;
    }
    p->cv_work.notify_all();
    return LSW_OK;
}

int lsw_pool_wait(lsw_pool* p, int64_t ticket, int64_t* n_hits, lsw_counters* counters, lsw_pool_times* times) {
    if (!p) return LSW_E_ARG;
    std::unique_lock<std::mutex> g(p->mu);
    auto it = p->live.find(ticket);
    if (it == p->live.end()) { p->err = "lsw_pool_wait: unknown ticket"; return LSW_E_ARG; }
    std::shared_ptr<Batch> b = it->second;
    p->cv_done.wait(g, [&] { return b->left == 0; });
    p->live.erase(ticket);
    const int D = (int)p->devices.size();
    const int R = (int)b->nhits.size();
    int64_t total = 0;
    for (int r = 0; r < R; r++) total += b->nhits[r];
    if (n_hits) *n_hits = total;
    if (counters) {
        lsw_counters c;
        std::memset(&c, 0, sizeof c);
        for (int d = 0; d < R; d++) {
            const lsw_counters& s = b->cnt[d];
            c.tasks += s.tasks; c.cells += s.cells; c.top_tasks += s.top_tasks; c.top_cells += s.top_cells; c.traced += s.traced;
            c.hits += s.hits; c.rounds = std::max(c.rounds, s.rounds);
            c.ms_score += s.ms_score / D; c.ms_trace += s.ms_trace / D;          // device time per device (mean over the devices)
            c.ms_other += s.ms_other / D; c.ms_total += s.ms_total / D;
            c.ms_score_round0 += s.ms_score_round0 / D;
            c.score_launches += s.score_launches; c.trace_launches += s.trace_launches; c.other_launches += s.other_launches;
            c.steps += s.steps; c.fallbacks += s.fallbacks; c.computed_cells += s.computed_cells;
            c.computed_cells_round0 += s.computed_cells_round0;
            c.device_bytes = std::max(c.device_bytes, s.device_bytes); c.reuse_dropped += s.reuse_dropped;
        }
        *counters = c;
    }
    if (times) {
        std::memset(times, 0, sizeof *times);
        std::vector<double> up(D, 0.0), fd(D, 0.0), ft(D, 0.0);
        for (int r = 0; r < R; r++) {
            const int d = r % D;
            up[d] += b->t_up[r]; fd[d] += b->t_find[r]; ft[d] += b->t_fetch[r];        // per device: the sum over its ranges
            times->gather_wait_s = std::max(times->gather_wait_s, b->t_wait[r]);
            if (d < 8) { times->reads[d] += b->cut[r + 1] - b->cut[r]; times->hits[d] += b->nhits[r]; }
        }
        for (int d = 0; d < D; d++) {
            times->upload_s = std::max(times->upload_s, up[d]); times->find_s = std::max(times->find_s, fd[d]);
            times->fetch_s = std::max(times->fetch_s, ft[d]);
        }
        times->total_s = b->t_done - b->t_submit;
    }
    if (b->rc != LSW_OK) p->err = b->err;
    return b->rc;
}

}  // extern "C"
