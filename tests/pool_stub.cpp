// TEST INFRASTRUCTURE: a device-free stand-in for the per-context entry points lsw_pool.cpp calls (include/lsw.h), so that the
// pool's host logic -- the cut into read ranges, the per-device turnstiles, the gather into the caller's buffer, error and overflow
// paths -- runs on a box without a GPU (tests/test_pool_host.py links lamprey_b200/csrc/lsw_pool.cpp against this file).
// A "context" produces len(read) % 3 hits per read (start = 0, 1), sleeps a pseudo-random while in every phase, and writes what it did
// into a global log.  Never part of the product.
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../include/lsw.h"

struct lsw_ctx {
    int device = 0;
    int64_t base = 0;
    std::vector<int64_t> len;
    std::vector<lsw_hit> hits;
    int64_t cap_user = 0;
    int64_t batch_id = 0;
    std::string err;
    uint32_t rng = 1;
};

namespace {
struct Event { int device; int phase; int64_t batch_id; int64_t base; };     // phase 0 = upload begins, 1 = find begins, 2 = fetch begins
std::mutex g_mu;
std::vector<Event> g_log;
int g_max_sleep_us = 2000;
int g_fail_device = -1;             // lsw_find_all fails on this device
int64_t g_default_cap = 1 << 30;    // device-side hit capacity before lsw_set_hit_capacity
int g_busy[64][3];                  // phases running at this moment, per device
int g_overlap_seen[3];              // a phase ran twice at once on one device

void nap(lsw_ctx* c) {
    c->rng = c->rng * 1664525u + 1013904223u;
    if (g_max_sleep_us > 0) std::this_thread::sleep_for(std::chrono::microseconds((c->rng >> 8) % (uint32_t)g_max_sleep_us));
}
struct Phase {
    int d, ph;
    Phase(lsw_ctx* c, int phase, int64_t id) : d(c->device), ph(phase) {
        std::lock_guard<std::mutex> g(g_mu);
        g_log.push_back({c->device, phase, id, c->base});
        if (++g_busy[d & 63][ph] > 1) g_overlap_seen[ph]++;
    }
    ~Phase() { std::lock_guard<std::mutex> g(g_mu); --g_busy[d & 63][ph]; }
};
}  // namespace

extern "C" {

int lsw_create(int device, lsw_ctx** out) {
    if (device < 0 || device >= 64) return LSW_E_CUDA;
    *out = new lsw_ctx();
    (*out)->device = device;
    (*out)->rng = 77u + (uint32_t)device;
    return LSW_OK;
}
void lsw_destroy(lsw_ctx* c) { delete c; }
const char* lsw_last_error(const lsw_ctx* c) { return c ? c->err.c_str() : "stub: no such device"; }
int lsw_set_scoring(lsw_ctx*, int, int, int, int) { return LSW_OK; }
int lsw_set_ssw(lsw_ctx*, int, int, int, int) { return LSW_OK; }
int lsw_set_queries(lsw_ctx* c, int n, const uint8_t*, const int32_t*) {
    if (n <= 0) { c->err = "stub: no queries"; return LSW_E_ARG; }
    return LSW_OK;
}
int lsw_set_overlap(lsw_ctx*, int) { return LSW_OK; }
int lsw_set_read_base(lsw_ctx* c, int64_t base) { c->base = base; return LSW_OK; }
int lsw_set_hit_capacity(lsw_ctx* c, int64_t cap) { c->cap_user = cap; return LSW_OK; }

int lsw_upload_reads(lsw_ctx* c, int64_t n, const uint8_t* ascii, const int64_t* offs) {
    // the batch is recognised by the first byte of its text (the pool hands every range the whole batch's text and the range's
    // own slice of the offsets): the tests write the batch number there
    c->batch_id = ascii && n > 0 && offs[n] > 0 ? (int64_t)ascii[0] : -1;
    Phase ph(c, 0, c->batch_id);
    nap(c);
    c->len.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) c->len[(size_t)i] = offs[i + 1] - offs[i];
    return LSW_OK;
}

int lsw_find_all(lsw_ctx* c, double, int64_t* n_hits, lsw_counters* cnt) {
    Phase ph(c, 1, c->batch_id);
    nap(c);
    if (c->device == g_fail_device) { c->err = "stub: find_all fails on this device"; return LSW_E_CUDA; }
    c->hits.clear();
    int64_t cells = 0;
    for (size_t i = 0; i < c->len.size(); i++) {
        for (int k = 0; k < (int)(c->len[i] % 3); k++) {
            lsw_hit h;
            std::memset(&h, 0, sizeof h);
            h.read = (int32_t)(c->base + (int64_t)i);
            h.start = k; h.end = k + 20; h.query = (int16_t)(c->len[i] % 7); h.matches = 20; h.score = 40; h.q_end = 20;
            c->hits.push_back(h);
        }
        cells += c->len[i];
    }
    if (cnt) {
        std::memset(cnt, 0, sizeof *cnt);
        cnt->tasks = (int64_t)c->len.size(); cnt->cells = cells; cnt->hits = (int64_t)c->hits.size(); cnt->rounds = 1;
    }
    *n_hits = (int64_t)c->hits.size();
    const int64_t cap = c->cap_user > 0 ? c->cap_user : g_default_cap;
    if (*n_hits > cap) { c->err = "stub: device hit buffer too small"; return LSW_E_OVERFLOW; }
    return LSW_OK;
}

int lsw_fetch_hits(lsw_ctx* c, lsw_hit* out, int64_t cap, int64_t* n_out) {
    Phase ph(c, 2, c->batch_id);
    nap(c);
    *n_out = (int64_t)c->hits.size();
    if (*n_out > cap) return LSW_E_OVERFLOW;
    if (*n_out) std::memcpy(out, c->hits.data(), c->hits.size() * sizeof(lsw_hit));
    return LSW_OK;
}

int lsw_fetch_hits_compact(lsw_ctx* c, lsw_hit16* out, int64_t cap, int64_t* n_out) {
    Phase ph(c, 2, c->batch_id);
    nap(c);
    *n_out = (int64_t)c->hits.size();
    if (*n_out > cap) return LSW_E_OVERFLOW;
    for (size_t i = 0; i < c->hits.size(); i++) {
        const lsw_hit& h = c->hits[i];
        lsw_hit16 o;
        std::memset(&o, 0, sizeof o);
        o.read = (uint32_t)h.read; o.start = (uint32_t)h.start; o.span = (uint8_t)(h.end - h.start); o.query = (uint8_t)h.query;
        o.matches = (uint8_t)h.matches; o.q_end = (uint8_t)h.q_end;
        out[i] = o;
    }
    return LSW_OK;
}

// ---- test controls ----
void stub_reset(int max_sleep_us, int fail_device, int64_t default_cap) {
    std::lock_guard<std::mutex> g(g_mu);
    g_log.clear();
    g_max_sleep_us = max_sleep_us; g_fail_device = fail_device; g_default_cap = default_cap;
    std::memset(g_busy, 0, sizeof g_busy);
    std::memset(g_overlap_seen, 0, sizeof g_overlap_seen);
}
int64_t stub_log(int64_t* out /* [cap][4] */, int64_t cap) {
    std::lock_guard<std::mutex> g(g_mu);
    const int64_t n = (int64_t)g_log.size();
    for (int64_t i = 0; i < n && i < cap; i++) {
        out[4 * i] = g_log[(size_t)i].device; out[4 * i + 1] = g_log[(size_t)i].phase;
        out[4 * i + 2] = g_log[(size_t)i].batch_id; out[4 * i + 3] = g_log[(size_t)i].base;
    }
    return n;
}
void stub_overlaps(int* out /* [3] */) {
    std::lock_guard<std::mutex> g(g_mu);
    for (int k = 0; k < 3; k++) out[k] = g_overlap_seen[k];
}

}  // extern "C"
