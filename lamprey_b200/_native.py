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
                        ("steps", "<i8"), ("fallbacks", "<i8"), ("computed_cells", "<i8"), ("device_bytes", "<i8"),
                        ("reuse_dropped", "<i8"), ("ms_score_round0", "<f4"), ("reserved3", "<f4"),
                        ("computed_cells_round0", "<i8")])
# lsw_hit16: the same hit in 16 bytes (include/lsw.h)
HIT16_DT = np.dtype([("read", "<u4"), ("start", "<u4"), ("span", "u1"), ("query", "u1"), ("depth", "<u2"),
                     ("matches", "u1"), ("mismatches", "u1"), ("q_pos", "u1"), ("q_end", "u1")])
assert TASK_DT.itemsize == 16 and RESULT_DT.itemsize == 40 and HIT_DT.itemsize == 32 and COUNTERS_DT.itemsize == 152
assert HIT16_DT.itemsize == 16

# every symbol include/lsw.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = ("lsw_create", "lsw_destroy", "lsw_last_error", "lsw_set_stream", "lsw_set_scoring", "lsw_set_queries",
           "lsw_upload_reads", "lsw_align_tasks", "lsw_set_read_base", "lsw_set_overlap", "lsw_set_hit_capacity", "lsw_find_all", "lsw_fetch_hits",
           "lsw_measure_int_peak", "lsw_version", "lsw_fetch_hits_compact", "lsw_set_ssw", "lsw_chain_subreads", "lsw_ssw_subreads", "lsw_host_alloc", "lsw_host_free",
           "lsw_pool_create", "lsw_pool_destroy", "lsw_pool_last_error", "lsw_pool_configure", "lsw_pool_configure_ssw", "lsw_pool_set_pieces", "lsw_pool_submit", "lsw_pool_wait",
           "lsw_fastq_index", "lsw_fastq_extract", "lsw_fastq_free", "lsw_expand_hits", "lsw_hits_index", "lsw_host_last_error")
PAIR_DT = np.dtype([("read", "<i4"), ("start", "<i4"), ("end", "<i4"), ("revcomp", "<i4")])
SUBREAD_DT = np.dtype([("read", "<i4"), ("start", "<i4"), ("end", "<i4"), ("target_idx", "<i4"), ("direction", "<i4"), ("hit", "<i4")])
HITS_WIDE, HITS_COMPACT = 0, 1
POOL_TIMES_DT = np.dtype([("upload_s", "<f8"), ("find_s", "<f8"), ("fetch_s", "<f8"), ("gather_wait_s", "<f8"), ("total_s", "<f8"),
                          ("reads", "<i8", (8,)), ("hits", "<i8", (8,))])

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
    lib.lsw_fetch_hits_compact.argtypes = [vp, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_set_ssw.argtypes = [vp, i32, i32, i32, i32]
    lib.lsw_ssw_subreads.argtypes = [vp, i64, vp, vp, i32, vp, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_chain_subreads.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_host_alloc.argtypes = [i64, ctypes.POINTER(vp)]
    lib.lsw_host_free.argtypes = [vp]
    lib.lsw_host_free.restype = None
    lib.lsw_pool_create.argtypes = [i32, ctypes.POINTER(i32), i32, ctypes.POINTER(vp)]
    lib.lsw_pool_destroy.argtypes = [vp]
    lib.lsw_pool_destroy.restype = None
    lib.lsw_pool_last_error.argtypes = [vp]
    lib.lsw_pool_last_error.restype = ctypes.c_char_p
    lib.lsw_pool_configure.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, dbl, i32, i32]
    lib.lsw_pool_configure_ssw.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, dbl, i32, i32]
    lib.lsw_pool_set_pieces.argtypes = [vp, i32]
    lib.lsw_pool_submit.argtypes = [vp, i64, vp, vp, vp, i64, ctypes.POINTER(i64)]
    lib.lsw_pool_wait.argtypes = [vp, i64, ctypes.POINTER(i64), vp, vp]
    lib.lsw_fastq_index.argtypes = [vp, i64, i32, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(i64)]
    lib.lsw_fastq_extract.argtypes = [vp, vp, vp, vp]
    lib.lsw_fastq_free.argtypes = [vp]
    lib.lsw_fastq_free.restype = None
    lib.lsw_expand_hits.argtypes = [vp, i64, i32, i32, i32, vp, i32]
    lib.lsw_hits_index.argtypes = [vp, i64, i32, i64, i64, vp, vp, i32]
    lib.lsw_host_last_error.argtypes = []
    lib.lsw_host_last_error.restype = ctypes.c_char_p
    for name in SYMBOLS:
        if name not in ("lsw_destroy", "lsw_last_error", "lsw_host_free", "lsw_pool_destroy", "lsw_pool_last_error", "lsw_fastq_free",
                        "lsw_host_last_error"):
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


def _host_check(lib, rc, what):
    if rc != LSW_OK:
        raise ValueError("%s failed (%d): %s" % (what, rc, lib.lsw_host_last_error().decode()))


def expand_hits(h16, match=2, mismatch=-1, gap=-1, threads=0):
    """lsw_hit16 records -> HIT_DT records (end, score and the survivor flag are re-derived: see include/lsw.h); lsw_expand_hits,
    all host cores."""
    lib = load()
    h16 = np.ascontiguousarray(h16, HIT16_DT)
    out = np.empty(len(h16), HIT_DT)
    _host_check(lib, lib.lsw_expand_hits(_ptr(h16), len(h16), int(match), int(mismatch), int(gap), _ptr(out), int(threads)), "lsw_expand_hits")
    return out


def hits_index(hits, n_reads, read_base=0, threads=0):
    """-> (offsets int64[n_reads + 1], covered int64[n_reads]) of a hit list sorted by read, in either record format
    (lsw_hits_index): read r's hits are hits[offsets[r]:offsets[r + 1]], covered[r] = sum of (end - start) over them."""
    lib = load()
    if hits.dtype == HIT16_DT:
        fmt = HITS_COMPACT
    elif hits.dtype == HIT_DT:
        fmt = HITS_WIDE
    else:
        raise TypeError("hits must be HIT_DT or HIT16_DT records")
    hits = np.ascontiguousarray(hits)
    offsets = np.empty(int(n_reads) + 1, np.int64)
    covered = np.empty(int(n_reads), np.int64)
    _host_check(lib, lib.lsw_hits_index(_ptr(hits), len(hits), fmt, int(n_reads), int(read_base), ctypes.c_void_p(offsets.ctypes.data),
                                        _ptr(covered), int(threads)), "lsw_hits_index")
    return offsets, covered


class PinnedArray(object):
    """A numpy array over page-locked host memory (lsw_host_alloc): copies from / to it run at PCIe speed.  Keep the
    object alive while the array is in use; close() (or garbage collection) frees the memory."""

    def __init__(self, count, dtype):
        lib = load()
        self.dtype = np.dtype(dtype)
        nbytes = max(1, int(count) * self.dtype.itemsize)
        p = ctypes.c_void_p(0)
        rc = lib.lsw_host_alloc(nbytes, ctypes.byref(p))
        if rc != LSW_OK:
            raise MemoryError("lsw_host_alloc(%d) failed (%d)" % (nbytes, rc))
        self._lib, self._p = lib, p
        buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=int(count) * self.dtype.itemsize).view(self.dtype)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self._lib.lsw_host_free(self._p)
            self._p = None

    __del__ = close


class Pool(object):
    """Several contexts behind one handle (include/lsw.h "Pool"): one batch is split over `devices` by bases and the hits come
    back gathered in global read order; with contexts_per_device = 2 consecutive batches overlap copies and kernels."""

    def __init__(self, devices=(0,), contexts_per_device=1):
        lib = load()
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = ctypes.c_void_p(0)
        rc = lib.lsw_pool_create(len(devices), devs, int(contexts_per_device), ctypes.byref(h))
        if rc != LSW_OK:
            raise RuntimeError("lsw_pool_create failed (%d): %s" % (rc, lib.lsw_pool_last_error(None).decode()))
        self._lib, self._h = lib, h
        self.devices = [int(d) for d in devices]
        self.contexts_per_device = int(contexts_per_device)
        self.hit_format = HITS_WIDE
        self.scoring = (2, -1, -1)
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lsw_pool_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc, what):
        if rc == LSW_OK:
            return
        msg = "%s failed (%d): %s" % (what, rc, self._lib.lsw_pool_last_error(self._h).decode())
        if rc == LSW_E_UNSUPPORTED:
            raise NotImplementedError(msg) if "not implemented" in msg else ValueError(msg)
        if rc in (LSW_E_ARG, LSW_E_STATE):
            raise ValueError(msg)
        if rc == LSW_E_OVERFLOW:
            raise OverflowError(msg)
        raise RuntimeError(msg)

    def configure(self, seqs, threshold, match=2, mismatch=-1, gap=-1, gap_ext=-1, allowed_overlap=None, compact=False, pieces=1,
                  ssw=None):
        """ssw=(match, mismatch, gap_open, gap_extend): the default (skbio) seeding branch instead of the swalign one."""
        concat, offs = concat_sequences(seqs)
        offs32 = offs.astype(np.int32)
        fmt = HITS_COMPACT if compact else HITS_WIDE
        ov = -1 if allowed_overlap is None else int(allowed_overlap)
        if ssw is not None:
            self._check(self._lib.lsw_pool_configure_ssw(self._h, int(ssw[0]), int(ssw[1]), int(ssw[2]), int(ssw[3]), len(seqs), _ptr(concat),
                                                         _ptr(offs32), float(threshold), ov, fmt), "lsw_pool_configure_ssw")
        else:
            self._check(self._lib.lsw_pool_configure(self._h, int(match), int(mismatch), int(gap), int(gap_ext), len(seqs), _ptr(concat),
                                                     _ptr(offs32), float(threshold), ov, fmt), "lsw_pool_configure")
        self._check(self._lib.lsw_pool_set_pieces(self._h, int(pieces)), "lsw_pool_set_pieces")
        self.hit_format = fmt
        self.scoring = (int(match), int(mismatch), int(gap))

    @property
    def hit_dtype(self):
        return HIT16_DT if self.hit_format == HITS_COMPACT else HIT_DT

    def submit(self, ascii_u8, offs_i64, out):
        """-> ticket.  `out` is a hit_dtype array the hits are written into; the three arrays must stay alive (and
        unmodified) until wait(ticket)."""
        a = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        o = np.ascontiguousarray(offs_i64, dtype=np.int64)
        assert out.dtype == self.hit_dtype and out.flags["C_CONTIGUOUS"]
        t = ctypes.c_int64(0)
        self._check(self._lib.lsw_pool_submit(self._h, len(o) - 1, _ptr(a), _ptr(o), _ptr(out), len(out), ctypes.byref(t)),
                    "lsw_pool_submit")
        self._keep[t.value] = (a, o, out)
        return t.value

    def wait(self, ticket):
        """-> (hits view of the submitted buffer, counters dict, times dict).  OverflowError carries the needed count in
        .needed when the buffer was too small."""
        n = ctypes.c_int64(0)
        cnt = np.zeros(1, COUNTERS_DT)
        tm = np.zeros(1, POOL_TIMES_DT)
        rc = self._lib.lsw_pool_wait(self._h, int(ticket), ctypes.byref(n), _ptr(cnt), _ptr(tm))
        _, _, out = self._keep.pop(ticket, (None, None, None))
        if rc == LSW_E_OVERFLOW:
            e = OverflowError("hit buffer too small: %d records are needed" % n.value)
            e.needed = n.value
            raise e
        self._check(rc, "lsw_pool_wait")
        counters = {k: cnt[0][k].item() for k in COUNTERS_DT.names}
        times = {k: (tm[0][k].tolist() if tm[0][k].ndim else tm[0][k].item()) for k in POOL_TIMES_DT.names}
        return out[:n.value], counters, times


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

    def set_ssw(self, match=2, mismatch=-3, gap_open=5, gap_extend=2):
        """The default seeding path: skbio StripedSmithWaterman semantics (include/lsw.h lsw_set_ssw)."""
        key = ("ssw", int(match), int(mismatch), int(gap_open), int(gap_extend))
        if key == self._scoring:
            return
        self._scoring = None
        self._check(self._lib.lsw_set_ssw(self._h, *key[1:]), "lsw_set_ssw")
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

    def ssw_subreads(self, pairs, target):
        """Stage 2 of processLamplicon: StripedSmithWaterman(sub-read)(target) for PAIR_DT sub-reads of the uploaded batch.
        -> (results RESULT_DT, cigar uint32[])."""
        pairs = np.ascontiguousarray(pairs, dtype=PAIR_DT)
        t = np.frombuffer(target.encode("ascii") if isinstance(target, str) else bytes(target), dtype=np.uint8)
        n = len(pairs)
        out = np.zeros(n, RESULT_DT)
        used = ctypes.c_int64(0)
        cap = max(64, 16 * n)
        while True:
            cig = np.zeros(cap, np.uint32)
            rc = self._lib.lsw_ssw_subreads(self._h, n, _ptr(pairs), _ptr(t), len(t), _ptr(out), _ptr(cig), cap, ctypes.byref(used))
            if rc == LSW_E_OVERFLOW and used.value > cap:
                cap = used.value
                continue
            self._check(rc, "lsw_ssw_subreads")
            return out, cig[:used.value]

    def chain_subreads(self, fwd_order, rev_order, t_query, tc_query):
        """Stage 1 of processLamplicon for the last find_all (which must have run with set_overlap): SUBREAD_DT records."""
        fo = np.ascontiguousarray(fwd_order, dtype=np.int32)
        ro = np.ascontiguousarray(rev_order, dtype=np.int32)
        cap = 1024
        while True:
            out = np.zeros(cap, SUBREAD_DT)
            n = ctypes.c_int64(0)
            rc = self._lib.lsw_chain_subreads(self._h, len(fo), _ptr(fo), len(ro), _ptr(ro), int(t_query), int(tc_query), _ptr(out), cap,
                                              ctypes.byref(n))
            if rc == LSW_E_OVERFLOW and n.value > cap:
                cap = n.value
                continue
            self._check(rc, "lsw_chain_subreads")
            return out[:n.value]

    def fetch_hits_compact(self, n=None, out=None):
        """The hits of the last find_all as 16-byte records (HIT16_DT); expand_hits() turns them into HIT_DT."""
        if out is None:
            out = np.zeros(int(n), HIT16_DT)
        got = ctypes.c_int64(0)
        self._check(self._lib.lsw_fetch_hits_compact(self._h, _ptr(out), len(out), ctypes.byref(got)), "lsw_fetch_hits_compact")
        return out[:got.value]

    def measure_int_peak(self, mode, iters=20000):
        ops = ctypes.c_double(0)
        ms = ctypes.c_float(0)
        self._check(self._lib.lsw_measure_int_peak(self._h, int(mode), int(iters), ctypes.byref(ops), ctypes.byref(ms)),
                    "lsw_measure_int_peak")
        return ops.value, ms.value
