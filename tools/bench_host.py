"""Host side of a FASTQ -> hits run, measured without a GPU: the FASTQ scanner (lsw_fastq_*), the 16 -> 32 byte record expansion
(lsw_expand_hits) and the per-read index + coverage of a hit list (lsw_hits_index), each beside the numpy code it replaced.
    python tools/bench_host.py [reads]      (run from the repo root; prints one JSON line)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lamprey_b200 import _native, fastq, seeding, synth   # noqa: E402


def best(f, n=3):
    out = []
    for _ in range(n):
        t = time.perf_counter(); f(); out.append(time.perf_counter() - t)
    return min(out)


def numpy_fastq(path):               # the reader this repo used before (round 1)
    raw = np.fromfile(path, dtype=np.uint8)
    nl = np.flatnonzero(raw == 10)
    starts = np.concatenate(([0], nl[:-1] + 1))
    s, e = starts[1::4], nl[1::4]
    lens = (e - s).astype(np.int64)
    offs = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    return raw[np.repeat(s - offs[:-1], lens) + np.arange(offs[-1], dtype=np.int64)], offs


def numpy_expand(h16):
    out = np.zeros(len(h16), _native.HIT_DT)
    m = (h16["matches"] & 0x7F).astype(np.int32); x = h16["mismatches"].astype(np.int32); span = h16["span"].astype(np.int32)
    qspan = h16["q_end"].astype(np.int32) - h16["q_pos"]
    a = span + qspan - m - x
    out["read"] = h16["read"]; out["start"] = h16["start"]; out["end"] = h16["start"].astype(np.int64) + span
    out["query"] = h16["query"]; out["depth"] = np.minimum(h16["depth"], 32767); out["matches"] = m; out["mismatches"] = x
    out["score"] = 2 * m - (a - m) - ((span - a) + (qspan - a)); out["q_pos"] = h16["q_pos"]; out["q_end"] = h16["q_end"]
    out["reserved"] = h16["matches"] >> 7
    return out


def numpy_index(h, n_reads):
    offs = np.searchsorted(h["read"], np.arange(n_reads + 1))
    cov = np.zeros(n_reads, np.int64)
    np.add.at(cov, h["read"], (h["end"] - h["start"]).astype(np.int64))
    return offs, cov


def measure(concat, offs, with_numpy=True, per=128):
    """-> dict of timings for the reads (concat, offs): FASTQ scan, record expansion, hit-list index.  per: synthetic hits per read
    (128 = what a 2 kb concatemer read of the headline workload yields)."""
    n = len(offs) - 1
    path = os.path.join(tempfile.gettempdir(), "bench_host_%d.fastq" % os.getpid())
    b = concat.tobytes()
    try:
        with open(path, "wb") as fp:
            for i in range(n):
                L = offs[i + 1] - offs[i]
                fp.write(b"@read%d runid=abc ch=%d\n" % (i, i)); fp.write(b[offs[i]:offs[i + 1]]); fp.write(b"\n+\n"); fp.write(b"I" * L); fp.write(b"\n")
        size = os.path.getsize(path)
        got = fastq.read_fastq(path)
        assert (got[0] == concat).all() and (got[1] == offs).all()
        t_new = best(lambda: fastq.read_fastq(path))
        t_old = best(lambda: numpy_fastq(path), 1) if with_numpy else None
    finally:
        if os.path.exists(path):
            os.remove(path)
    N = n * per
    rng = np.random.default_rng(0)
    h = np.zeros(N, _native.HIT16_DT)
    h["read"] = np.repeat(np.arange(n, dtype=np.uint32), per); h["start"] = np.tile(np.arange(per, dtype=np.uint32) * 15, n)
    h["span"] = 24; h["query"] = rng.integers(0, 46, N); h["matches"] = 20; h["mismatches"] = 3; h["q_end"] = 23
    w = _native.expand_hits(h)
    if with_numpy:
        assert w.tobytes() == numpy_expand(h).tobytes()
    e_new, e_old = best(lambda: _native.expand_hits(h)), (best(lambda: numpy_expand(h), 1) if with_numpy else None)
    i_new, i_old = best(lambda: _native.hits_index(h, n)), (best(lambda: numpy_index(w, n), 1) if with_numpy else None)
    names = ["q%d" % i for i in range(46)]
    sb = best(lambda: seeding.SeedBatch(w, n, np.diff(offs), names, {}))
    sp = (lambda old, new: old / new if old else None)
    return {"cores": os.cpu_count(), "reads": n, "fastq_bytes": size, "hits": N,
            "fastq": {"native_s": t_new, "numpy_s": t_old, "native_reads_per_s": n / t_new, "native_gb_per_s": size / t_new / 1e9,
                      "speedup": sp(t_old, t_new)},
            "expand_hits": {"native_s": e_new, "numpy_s": e_old, "native_mhits_per_s": N / e_new / 1e6, "speedup": sp(e_old, e_new)},
            "hits_index": {"native_s": i_new, "numpy_s": i_old, "native_mhits_per_s": N / i_new / 1e6, "speedup": sp(i_old, i_new)},
            "seedbatch_construct_s": sb}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    concat, offs = synth.generate(n, workers=os.cpu_count(), **synth.CONFIGS["synth_2kb_8pct"])
    print(json.dumps(measure(concat, offs)))


if __name__ == "__main__":
    main()
