// lsw_hostops.cpp -- the host-side work on either side of the device path, multi-threaded (include/lsw.h, "Host side").
//
// Once seeding runs at ~2 * 10^6 reads/s per GPU, what is left on the host decides the speed of a FASTQ -> hits run:
//   * FASTQ ingest.  The reference reads its input with Biopython (SeqIO.parse(fn, "fastq"), /root/reference/lamprey.py:2145-2155;
//     per delivered file in online mode, lamprey.py:451-507) and keeps str(record.seq) of every record.  lsw_fastq_index /
//     lsw_fastq_extract turn the bytes of a FASTQ file straight into the (text, offsets) pair lsw_upload_reads / lsw_pool_submit
//     take -- no per-read objects.  The file is cut into one piece per thread; because records are exactly four lines, the number of
//     newlines before a piece tells every thread which line of a record it starts on.
//   * what findAllPrimerAlignments returns next to the hits (lamprey.py:722-731): the coverage of every read, and where a read's hits
//     start in the sorted list (lsw_hits_index); and the 32-byte view of 16-byte hit records (lsw_expand_hits).
// No CUDA in this file.
#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lsw.h"

namespace {

thread_local std::string t_host_error;

int host_fail(int rc, const std::string& msg) { t_host_error = msg; return rc; }

int pick_threads(int threads, int64_t work_items, int64_t min_per_thread) {
    int t = threads > 0 ? threads : (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
    const int64_t by_work = std::max<int64_t>(1, work_items / std::max<int64_t>(1, min_per_thread));
    return (int)std::max<int64_t>(1, std::min<int64_t>(t, by_work));
}

template <class F> void parallel_for(int T, F f) {
    if (T <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)T - 1);
    for (int t = 1; t < T; t++) th.emplace_back(f, t);
    f(0);
    for (auto& x : th) x.join();
}

int64_t count_newlines(const uint8_t* p, const uint8_t* e) {
    int64_t n = 0;
    while (p < e) {
        const void* q = std::memchr(p, '\n', (size_t)(e - p));
        if (!q) break;
        n++;
        p = static_cast<const uint8_t*>(q) + 1;
    }
    return n;
}

struct Rec { int64_t hdr, hdr_end, seq; int32_t seq_len, qual_len; };

}  // namespace

struct lsw_fastq {
    const uint8_t* data = nullptr;
    int64_t size = 0;                 // without trailing newlines
    std::vector<Rec> rec;
    int64_t n_bases = 0;
    int threads = 1;
};

extern "C" {

const char* lsw_host_last_error(void) { return t_host_error.c_str(); }

int lsw_fastq_index(const uint8_t* data, int64_t size, int threads, lsw_fastq** out, int64_t* n_reads, int64_t* n_bases) {
    if (!out || size < 0 || (size > 0 && !data)) return host_fail(LSW_E_ARG, "lsw_fastq_index: bad argument");
    *out = nullptr;
    while (size > 0 && (data[size - 1] == '\n' || data[size - 1] == '\r')) size--;       // blank lines at the end of the file
    lsw_fastq* fq = new lsw_fastq();
    fq->data = data; fq->size = size;
    if (size == 0) {
        if (n_reads) *n_reads = 0;
        if (n_bases) *n_bases = 0;
        *out = fq;
        return LSW_OK;
    }
    const int T = pick_threads(threads, size, 1 << 20);
    fq->threads = T;
    std::vector<int64_t> cut((size_t)T + 1), nl((size_t)T + 1, 0);
    for (int t = 0; t <= T; t++) cut[(size_t)t] = size * t / T;
    parallel_for(T, [&](int t) { nl[(size_t)t + 1] = count_newlines(data + cut[(size_t)t], data + cut[(size_t)t + 1]); });
    for (int t = 0; t < T; t++) nl[(size_t)t + 1] += nl[(size_t)t];
    int64_t lines = nl[(size_t)T] + 1;                               // the last line has no newline any more
    if (lines % 4 == 3) {
        // a last record with an EMPTY read: its empty quality line went with the blank lines at the end of the file
        const void* q = memrchr(data, '\n', (size_t)size);
        const int64_t ls = q ? (static_cast<const uint8_t*>(q) - data) + 1 : 0;
        if (ls < size && data[ls] == '+') lines++;                    // (a non-empty read without quality values fails the length check)
    }
    if (lines % 4) {
        delete fq;
        return host_fail(LSW_E_ARG, "FASTQ: " + std::to_string(lines) + " lines is not a multiple of 4 (multi-line records are not supported)");
    }
    const int64_t n = lines / 4;
    if (n >= (1ll << 31)) { delete fq; return host_fail(LSW_E_ARG, "FASTQ: at most 2^31 - 1 records per file"); }
    fq->rec.resize((size_t)n);
    std::vector<int64_t> bad((size_t)T, -1);                         // first malformed line seen by a thread
    std::vector<int> why((size_t)T, 0);
    parallel_for(T, [&](int t) {
        const int64_t lo = cut[(size_t)t], hi = cut[(size_t)t + 1];
        // the first line that STARTS in [lo, hi), and its number
        int64_t p = lo, line = nl[(size_t)t];
        if (lo > 0 && data[lo - 1] != '\n') {
            const void* q = std::memchr(data + lo, '\n', (size_t)(size - lo));
            if (!q) return;                                           // the piece lies inside the last line, which an earlier thread owns
            p = (static_cast<const uint8_t*>(q) - data) + 1;
            line = nl[(size_t)t] + 1;                                 // (no other newline between lo and p)
            if (p >= hi) return;                                      // that line starts in a later piece
        }
        while (p < hi) {
            const void* q = std::memchr(data + p, '\n', (size_t)(size - p));
            const int64_t next = q ? (static_cast<const uint8_t*>(q) - data) : size;
            int64_t end = next;
            if (end > p && data[end - 1] == '\r') end--;              // CRLF
            Rec& r = fq->rec[(size_t)(line >> 2)];
            switch (line & 3) {
            case 0:
                if (end <= p || data[p] != '@') { if (bad[(size_t)t] < 0) { bad[(size_t)t] = line; why[(size_t)t] = 0; } }
                r.hdr = p; r.hdr_end = end;
                break;
            case 1:
                if (end - p >= (1ll << 27)) { if (bad[(size_t)t] < 0) { bad[(size_t)t] = line; why[(size_t)t] = 2; } }
                r.seq = p; r.seq_len = (int32_t)std::min<int64_t>(end - p, (1ll << 27));
                break;
            case 2:
                if (end <= p || data[p] != '+') { if (bad[(size_t)t] < 0) { bad[(size_t)t] = line; why[(size_t)t] = 1; } }
                break;
            default:
                r.qual_len = (int32_t)std::min<int64_t>(end - p, (1ll << 27));
                break;
            }
            line++;
            p = next + 1;
        }
    });
    for (int t = 0; t < T; t++)
        if (bad[(size_t)t] >= 0) {
            const int64_t line = bad[(size_t)t];
            delete fq;
            const char* msg[3] = {"a record does not start with '@'", "the third line of a record does not start with '+'", "a read of 2^27 bases or more"};
            return host_fail(LSW_E_ARG, "FASTQ: not a 4-line FASTQ: line " + std::to_string(line + 1) + ": " + msg[why[(size_t)t]]);
        }
    // Biopython refuses records whose quality string and sequence differ in length
    const int T2 = pick_threads(threads, n, 1 << 16);
    std::vector<int64_t> bases((size_t)T2 + 1, 0), badq((size_t)T2, -1);
    parallel_for(T2, [&](int t) {
        const int64_t lo = n * t / T2, hi = n * (t + 1) / T2;
        int64_t s = 0;
        for (int64_t i = lo; i < hi; i++) {
            const Rec& r = fq->rec[(size_t)i];
            if (r.qual_len != r.seq_len && badq[(size_t)t] < 0) badq[(size_t)t] = i;
            s += r.seq_len;
        }
        bases[(size_t)t + 1] = s;
    });
    for (int t = 0; t < T2; t++)
        if (badq[(size_t)t] >= 0) {
            const int64_t i = badq[(size_t)t];
            delete fq;
            return host_fail(LSW_E_ARG, "FASTQ: record " + std::to_string(i + 1) + ": lengths of sequence and quality values differ");
        }
    for (int t = 0; t < T2; t++) fq->n_bases += bases[(size_t)t + 1];
    if (n_reads) *n_reads = n;
    if (n_bases) *n_bases = fq->n_bases;
    *out = fq;
    return LSW_OK;
}

int lsw_fastq_extract(const lsw_fastq* fq, uint8_t* ascii_out, int64_t* offs_out, int64_t* name_spans) {
    if (!fq || !offs_out || (fq->n_bases > 0 && !ascii_out)) return host_fail(LSW_E_ARG, "lsw_fastq_extract: bad argument");
    const int64_t n = (int64_t)fq->rec.size();
    offs_out[0] = 0;
    for (int64_t i = 0; i < n; i++) offs_out[i + 1] = offs_out[i] + fq->rec[(size_t)i].seq_len;
    const int T = pick_threads(fq->threads, fq->n_bases, 1 << 20);
    parallel_for(T, [&](int t) {
        // pieces balanced by bases
        const int64_t b_lo = fq->n_bases * t / T, b_hi = fq->n_bases * (t + 1) / T;
        const int64_t lo = std::lower_bound(offs_out, offs_out + n + 1, b_lo) - offs_out;
        const int64_t hi = t + 1 == T ? n : std::lower_bound(offs_out, offs_out + n + 1, b_hi) - offs_out;
        for (int64_t i = lo; i < hi; i++) {
            const Rec& r = fq->rec[(size_t)i];
            if (r.seq_len) std::memcpy(ascii_out + offs_out[i], fq->data + r.seq, (size_t)r.seq_len);
            if (name_spans) {
                // Biopython's record.id: the title line after '@' up to the first white space
                int64_t a = r.hdr + 1, b = a;
                while (b < r.hdr_end && fq->data[b] != ' ' && fq->data[b] != '\t') b++;
                name_spans[2 * i] = a; name_spans[2 * i + 1] = b;
            }
        }
    });
    return LSW_OK;
}

void lsw_fastq_free(lsw_fastq* fq) { delete fq; }

int lsw_expand_hits(const lsw_hit16* in, int64_t n, int match, int mismatch, int gap, lsw_hit* out, int threads) {
    if (n < 0 || (n > 0 && (!in || !out))) return host_fail(LSW_E_ARG, "lsw_expand_hits: bad argument");
    const int T = pick_threads(threads, n, 1 << 18);
    parallel_for(T, [&](int t) {
        const int64_t lo = n * t / T, hi = n * (t + 1) / T;
        for (int64_t i = lo; i < hi; i++) {
            const lsw_hit16 c = in[i];
            lsw_hit h;
            const int m = c.matches & 0x7F, x = c.mismatches, span = c.span, qspan = (int)c.q_end - (int)c.q_pos;
            const int a = span + qspan - m - x;                       // aligned columns (include/lsw.h)
            h.read = (int32_t)c.read; h.start = (int32_t)c.start; h.end = (int32_t)(c.start + (uint32_t)span);
            h.query = c.query; h.depth = (int16_t)std::min<int>(c.depth, 32767);
            h.matches = (int16_t)m; h.mismatches = (int16_t)x;
            h.score = (int16_t)(match * m + mismatch * (a - m) + gap * ((span - a) + (qspan - a)));
            h.q_pos = c.q_pos; h.q_end = c.q_end;
            h.reserved = (int16_t)(c.matches >> 7); h.reserved2 = 0;
            out[i] = h;
        }
    });
    return LSW_OK;
}

int lsw_hits_index(const void* hits, int64_t n, int hit_format, int64_t n_reads, int64_t read_base, int64_t* offsets, int64_t* covered,
                   int threads) {
    if (n < 0 || n_reads < 0 || !offsets || (n > 0 && !hits) || (hit_format != LSW_HITS_WIDE && hit_format != LSW_HITS_COMPACT))
        return host_fail(LSW_E_ARG, "lsw_hits_index: bad argument");
    const lsw_hit* w = hit_format == LSW_HITS_WIDE ? static_cast<const lsw_hit*>(hits) : nullptr;
    const lsw_hit16* c = hit_format == LSW_HITS_COMPACT ? static_cast<const lsw_hit16*>(hits) : nullptr;
    auto read_of = [&](int64_t i) -> int64_t { return (w ? (int64_t)w[i].read : (int64_t)c[i].read) - read_base; };
    auto span_of = [&](int64_t i) -> int64_t { return w ? (int64_t)w[i].end - w[i].start : (int64_t)c[i].span; };
    const int T = pick_threads(threads, n, 1 << 18);
    std::vector<int> unsorted((size_t)T, 0);
    if (covered) std::memset(covered, 0, sizeof(int64_t) * (size_t)n_reads);
    // every thread owns a contiguous piece of the hit list and the reads whose FIRST hit lies in it: offsets[r] for the reads r in
    // (read of the hit before the piece, read of the piece's last hit]
    parallel_for(T, [&](int t) {
        const int64_t lo = n * t / T, hi = n * (t + 1) / T;
        int64_t prev = lo > 0 ? read_of(lo - 1) : -1;
        if (prev < -1 || prev >= n_reads) { unsorted[(size_t)t] = 1; return; }
        for (int64_t i = lo; i < hi; i++) {
            const int64_t r = read_of(i);
            if (r < prev || r < 0 || r >= n_reads) { unsorted[(size_t)t] = 1; return; }
            for (int64_t k = prev + 1; k <= r; k++) offsets[k] = i;
            prev = r;
        }
    });
    for (int t = 0; t < T; t++)
        if (unsorted[(size_t)t]) return host_fail(LSW_E_ARG, "lsw_hits_index: hits are not sorted by read, or a read id is outside [read_base, read_base + n_reads)");
    const int64_t last = n > 0 ? read_of(n - 1) : -1;
    for (int64_t k = last + 1; k <= n_reads; k++) offsets[k] = n;
    if (covered) {
        const int T2 = pick_threads(threads, n_reads, 1 << 14);
        parallel_for(T2, [&](int t) {
            const int64_t lo = n_reads * t / T2, hi = n_reads * (t + 1) / T2;
            for (int64_t r = lo; r < hi; r++) {
                int64_t s = 0;
                for (int64_t i = offsets[r]; i < offsets[r + 1]; i++) s += span_of(i);
                covered[r] = s;
            }
        });
    }
    return LSW_OK;
}

}  // extern "C"
