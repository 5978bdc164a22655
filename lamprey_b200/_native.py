"""ctypes binding of liblsw.so (C ABI: include/lsw.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``lamprey_b200/csrc/Makefile``.
There is NO CPU fallback: if the library is missing or no CUDA device is present, every
entry point raises ``RuntimeError``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LSW_LIB") or os.path.join(_HERE, "liblsw.so")     # LSW_LIB: A/B builds during tuning

LSW_OK, LSW_E_CUDA, LSW_E_ARG, LSW_E_UNSUPPORTED, LSW_E_OVERFLOW, LSW_E_STATE, LSW_E_INTERNAL = 0, -1, -2, -3, -4, -5, -6
MAX_QUERY_LEN = 63

TASK_DT = np.dtype([("read", "<i4"), ("start", "<i4"), ("end", "<i4"), ("query", "<i4")])
RESULT_DT = np.dtype([("score", "<i4"), ("q_pos", "<i4"), ("q_end", "<i4"), ("r_pos", "<i4"), ("r_end", "<i4"),
                      ("matches", "<i4"), ("mismatches", "<i4"), ("n_cigar", "<i4"), ("cigar_off", "<i8")])
HIT_DT = np.dtype([("read", "<i4"), ("start", "<i4"), ("end", "<i4"), ("query", "<i2"), ("depth", "<i2"),
                   ("matches", "<i2"), ("mismatches", "<i2"), ("score", "<i2"), ("q_pos", "<i2"), ("q_end", "<i2"),
                   ("reserved", "<i2"), ("reserved2", "<i4")])
COUNTERS_DT = np.dtype([("tasks", "<i8"), ("cells", "<i8"), ("top_tasks", "<i8"), ("top_cells", "<i8"),
                        ("traced", "<i8"), ("hits", "<i8"), ("rounds", "<i4"), ("reserved", "<i4"),
                        ("ms_score", "<f4"), ("ms_trace", "<f4"), ("ms_other", "<f4"), ("ms_total", "<f4"),
                        ("score_launches", "<i8"), ("trace_launches", "<i8"), ("other_launches", "<i8"),
                        ("steps", "<i8"), ("fallbacks", "<i8"), ("computed_cells", "<i8")])
assert TASK_DT.itemsize == 16 and RESULT_DT.itemsize == 40 and HIT_DT.itemsize == 32 and COUNTERS_DT.itemsize == 120

# every symbol include/lsw.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = ("lsw_create", "lsw_destroy", "lsw_last_error", "lsw_set_stream", "lsw_set_scoring", "lsw_set_queries",
           "lsw_upload_reads", "lsw_align_tasks", "lsw_set_read_base", "lsw_set_overlap", "lsw_set_hit_capacity", "lsw_find_all", "lsw_fetch_hits",
           "lsw_measure_int_peak", "lsw_version")

_lib = None


def load():
    """Open liblsw.so (no CUDA call is made until a context is created)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblsw.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "This engine has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    lib.lsw_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.lsw_destroy.argtypes = [vp]
    lib.lsw_destroy.restype = None
    lib.lsw_last_error.argtypes = [vp]
    lib.lsw_last_error.restype = ctypes.c_char_p
    lib.lsw_set_stream.argtypes = [vp, vp]
    lib.lsw_set_scoring.argtypes = [vp, i32, i32, i32, i32]
    lib.lsw_set_queries.argtypes = [vp, i32, vp, vp]
    lib.lsw_upload_reads.argtypes = [vp, i64, vp, vp]
    lib.lsw_align_tasks.argtypes = [vp, i64, vp, vp, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_set_hit_capacity.argtypes = [vp, i64]
    lib.lsw_set_read_base.argtypes = [vp, i64]
    lib.lsw_set_overlap.argtypes = [vp, i32]
    lib.lsw_find_all.argtypes = [vp, dbl, ctypes.POINTER(i64), vp]
    lib.lsw_fetch_hits.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_measure_int_peak.argtypes = [vp, i32, i32, ctypes.POINTER(dbl), ctypes.POINTER(ctypes.c_float)]
    for name in SYMBOLS:
        if name not in ("lsw_destroy", "lsw_last_error"):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else ctypes.c_void_p(0)


def concat_sequences(seqs):
    """list of str/bytes -> (uint8 concat, int offsets[n+1])."""
    bs = [s.encode("ascii") if isinstance(s, str) else bytes(s) for s in seqs]
    concat = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    offs = np.zeros(len(bs) + 1, np.int64)
    if bs:
        offs[1:] = np.cumsum([len(b) for b in bs])
    return concat, offs


class Engine(object):
    """One liblsw context (one GPU, one CUDA stream, single-threaded)."""

    def __init__(self, device=0):
        lib = load()
        h = ctypes.c_void_p(0)
        rc = lib.lsw_create(int(device), ctypes.byref(h))
        if rc != LSW_OK:
            raise RuntimeError("lsw_create failed (%d): %s" % (rc, lib.lsw_last_error(None).decode()))
        self._lib = lib
        self._h = h
        self.device = int(device)
        self.n_queries = 0
        self.n_reads = 0
        self.last_counters = None
        self._scoring = None          # what the context currently holds: repeated set_* calls with the same values are
        self._queries = None          # free (LocalAlignment.align sets both on every call, ~65 times per read)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lsw_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc == LSW_OK:
            return
        msg = "%s failed (%d): %s" % (what, rc, self._lib.lsw_last_error(self._h).decode())
        if rc == LSW_E_UNSUPPORTED:
            raise NotImplementedError(msg) if "not implemented" in msg else ValueError(msg)
        if rc in (LSW_E_ARG, LSW_E_STATE):
            raise ValueError(msg)
        if rc == LSW_E_OVERFLOW:
            raise OverflowError(msg)
        raise RuntimeError(msg)

    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.lsw_set_stream(self._h, ctypes.c_void_p(int(cuda_stream_ptr or 0))), "lsw_set_stream")

    def set_scoring(self, match=2, mismatch=-1, gap=-1, gap_ext=-1):
        key = (int(match), int(mismatch), int(gap), int(gap_ext))
        if key == self._scoring:
            return
        self._scoring = None
        self._check(self._lib.lsw_set_scoring(self._h, *key), "lsw_set_scoring")
        self._scoring = key

    def set_queries(self, seqs):
        key = tuple(s if isinstance(s, str) else bytes(s).decode("ascii") for s in seqs)
        if key == self._queries:
            return
        self._queries = None
        concat, offs = concat_sequences(key)
        offs32 = offs.astype(np.int32)
        self._check(self._lib.lsw_set_queries(self._h, len(key), _ptr(concat), _ptr(offs32)), "lsw_set_queries")
        self.n_queries = len(key)
        self._queries = key

    def upload_reads(self, ascii_u8, offs_i64):
        a = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        o = np.ascontiguousarray(offs_i64, dtype=np.int64)
        self._check(self._lib.lsw_upload_reads(self._h, len(o) - 1, _ptr(a), _ptr(o)), "lsw_upload_reads")
        self.n_reads = len(o) - 1

    def set_read_base(self, base):
        """Added to every hit's read index (for callers that split one batch over contexts or GPUs)."""
        self._check(self._lib.lsw_set_read_base(self._h, int(base)), "lsw_set_read_base")

    def set_overlap(self, allowed_overlap):
        """>= 0: hits come back with reserved = 1 if they survive removeOverlappingPrimerAlignments; None/-1: off."""
        self._check(self._lib.lsw_set_overlap(self._h, -1 if allowed_overlap is None else int(allowed_overlap)), "lsw_set_overlap")

    def upload_read_list(self, reads):
        concat, offs = concat_sequences(reads)
        self.upload_reads(concat, offs)

    def align_tasks(self, tasks, want_cigar=True):
        """tasks: structured array TASK_DT.  -> (results RESULT_DT, cigar uint32[])."""
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DT)
        n = len(tasks)
        out = np.zeros(n, RESULT_DT)
        used = ctypes.c_int64(0)
        if not want_cigar:
            self._check(self._lib.lsw_align_tasks(self._h, n, _ptr(tasks), _ptr(out), ctypes.c_void_p(0), 0,
                                                  ctypes.byref(used)), "lsw_align_tasks")
            return out, np.zeros(0, np.uint32)
        cap = max(64, 8 * n)
        while True:
            cig = np.zeros(cap, np.uint32)
            rc = self._lib.lsw_align_tasks(self._h, n, _ptr(tasks), _ptr(out), _ptr(cig), cap, ctypes.byref(used))
            if rc == LSW_E_OVERFLOW and used.value > cap:
                cap = used.value
                continue
            self._check(rc, "lsw_align_tasks")
            return out, cig[:used.value]

    def find_all(self, threshold, fetch=True):
        """-> hits (HIT_DT, sorted by read, start, query) or the hit count if fetch is False."""
        n = ctypes.c_int64(0)
        cnt = np.zeros(1, COUNTERS_DT)
        rc = self._lib.lsw_find_all(self._h, float(threshold), ctypes.byref(n), _ptr(cnt))
        if rc == LSW_E_OVERFLOW:
            # retry once with room for the reported count, then go back to the default capacity: a later, larger batch
            # must not inherit this batch's number
            self._check(self._lib.lsw_set_hit_capacity(self._h, n.value + 1024), "lsw_set_hit_capacity")
            try:
                rc = self._lib.lsw_find_all(self._h, float(threshold), ctypes.byref(n), _ptr(cnt))
            finally:
                self._lib.lsw_set_hit_capacity(self._h, 0)
        self._check(rc, "lsw_find_all")
        self.last_counters = {k: cnt[0][k].item() for k in COUNTERS_DT.names}
        if not fetch:
            return n.value
        return self.fetch_hits(n.value)

    def fetch_hits(self, n=None, out=None):
        if out is None:
            out = np.zeros(int(n), HIT_DT)
        got = ctypes.c_int64(0)
        self._check(self._lib.lsw_fetch_hits(self._h, _ptr(out), len(out), ctypes.byref(got)), "lsw_fetch_hits")
        return out[:got.value]

    def measure_int_peak(self, mode, iters=20000):
        ops = ctypes.c_double(0)
        ms = ctypes.c_float(0)
        self._check(self._lib.lsw_measure_int_peak(self._h, int(mode), int(iters), ctypes.byref(ops), ctypes.byref(ms)),
                    "lsw_measure_int_peak")
        return ops.value, ms.value
