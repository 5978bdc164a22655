#!/bin/bash
# The host-only C++ of the library under the sanitizers (no GPU needed):
#   ThreadSanitizer  over the pool (lsw_pool.cpp against the device-free stub tests/pool_stub.cpp): batches in flight on 3 "devices";
#   AddressSanitizer over the FASTQ scanner / hit-list helpers (lsw_hostops.cpp) on ragged inputs and every thread count.
# Run from the repo root.
set -e
TMP=${TMPDIR:-/tmp}
SRC="lamprey_b200/csrc/lsw_pool.cpp lamprey_b200/csrc/lsw_hostops.cpp tests/pool_stub.cpp"
g++ -O1 -g -std=c++17 -shared -fPIC -pthread -fsanitize=thread -o $TMP/libhost_tsan.so $SRC
g++ -O1 -g -std=c++17 -shared -fPIC -pthread -fsanitize=address -fno-omit-frame-pointer -o $TMP/libhost_asan.so $SRC
cat > $TMP/sanitize_host.py <<'PY'
import ctypes, os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
lib = ctypes.CDLL(os.environ["HOST_LIB"])
lib.lsw_pool_last_error.restype = ctypes.c_char_p
lib.lsw_host_last_error.restype = ctypes.c_char_p
lib.stub_reset.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64]
import test_pool_host as T
rng = np.random.default_rng(0)
for devices, per_dev, pieces in (([0, 1, 2], 2, 3), ([0], 2, 1), ([0, 1], 1, 2)):
    lib.stub_reset(300, -1, 1 << 30)
    h = T.make_pool(lib, devices, per_dev, pieces, True)
    outs = [np.zeros(1000, T.HIT16_DT) for _ in range(4)]
    batches = [T.make_batch(k, int(rng.integers(40, 400)), rng) for k in range(16)]
    tickets = []
    for k, (text, offs) in enumerate(batches):
        if k >= 4:
            assert T.wait(lib, h, tickets[k - 4])[0] == 0
        tickets.append(T.submit(lib, h, text, offs, outs[k % 4])[1])
    for t in tickets[-4:]:
        assert T.wait(lib, h, t)[0] == 0
    lib.lsw_pool_destroy(h)
print("pool: done")
P = lambda a: ctypes.c_void_p(a.ctypes.data) if a.size else ctypes.c_void_p(0)
for threads in (1, 2, 7, 16):
    for n in (0, 1, 3, 500):
        recs = []
        for i in range(n):
            L = int(rng.choice([0, 1, 7, 300, 5000]))
            recs.append(b"@r%d x\n" % i + bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L)) + b"\n+\n" + b"I" * L + b"\n")
        for tail in (b"", b"\n\n"):
            data = np.frombuffer(b"".join(recs).rstrip(b"\n") + tail, np.uint8).copy()
            fq = ctypes.c_void_p(0); nr = ctypes.c_int64(0); nb = ctypes.c_int64(0)
            rc = lib.lsw_fastq_index(P(data), ctypes.c_int64(data.size), threads, ctypes.byref(fq), ctypes.byref(nr), ctypes.byref(nb))
            assert rc == 0 and nr.value == n, (rc, lib.lsw_host_last_error())
            text = np.empty(nb.value, np.uint8); offs = np.empty(n + 1, np.int64); spans = np.empty((n, 2), np.int64)
            assert lib.lsw_fastq_extract(fq, P(text), P(offs), P(spans)) == 0
            lib.lsw_fastq_free(fq)
    n_reads = 777
    counts = rng.integers(0, 9, n_reads)
    c = np.zeros(int(counts.sum()), T.HIT16_DT)
    c["read"] = np.repeat(np.arange(n_reads), counts); c["span"] = 20
    w = np.empty(len(c), T.HIT_DT)
    assert lib.lsw_expand_hits(P(c), ctypes.c_int64(len(c)), 2, -1, -1, P(w), threads) == 0
    offs = np.empty(n_reads + 1, np.int64); cov = np.empty(n_reads, np.int64)
    for hits, fmt in ((c, 1), (w, 0)):
        assert lib.lsw_hits_index(P(hits), ctypes.c_int64(len(hits)), fmt, ctypes.c_int64(n_reads), ctypes.c_int64(0), P(offs), P(cov), threads) == 0
        assert cov.sum() == 20 * len(c)
print("host ops: done")
PY
echo "== ThreadSanitizer"
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="report_bugs=1 exitcode=66" HOST_LIB=$TMP/libhost_tsan.so python $TMP/sanitize_host.py
echo "== AddressSanitizer"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 HOST_LIB=$TMP/libhost_asan.so python $TMP/sanitize_host.py
echo "sanitizers: clean"
