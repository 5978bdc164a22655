"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/libsw_oracle.so
(the C restatement in sw_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libsw_oracle.so")
_lib = None

RESULT_DT = np.dtype([("score", "<i4"), ("q_pos", "<i4"), ("q_end", "<i4"), ("r_pos", "<i4"),
                      ("r_end", "<i4"), ("matches", "<i4"), ("mismatches", "<i4"), ("n_cigar", "<i4")])
HIT_DT = np.dtype([("seq", "<i4"), ("start", "<i4"), ("end", "<i4"), ("matches", "<i4"),
                   ("mismatches", "<i4"), ("score", "<i4"), ("q_pos", "<i4"), ("q_end", "<i4"),
                   ("depth", "<i4")])
STATS_DT = np.dtype([("calls", "<i8"), ("cells", "<i8"), ("max_depth", "<i4"), ("_pad", "<i4")])
_OPS = " MID"


def build(force=False):
    src = os.path.join(HERE, "sw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.swo_align.restype = ctypes.c_int
        _lib.swo_find_all.restype = ctypes.c_int
    return _lib


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray)) else np.ascontiguousarray(b, np.uint8)
    return a


def align(ref, query, match=2, mismatch=-1, gap=-1, gap_ext=-1):
    """-> dict(score,q_pos,q_end,r_pos,r_end,matches,mismatches,cigar=[(n,op)])."""
    r = _buf(ref.encode() if isinstance(ref, str) else ref)
    q = _buf(query.encode() if isinstance(query, str) else query)
    out = np.zeros(1, RESULT_DT)
    cap = len(r) + len(q) + 2
    cig = np.zeros(cap, np.uint32)
    rc_ = lib().swo_align(r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(r)),
                          q.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(q)),
                          ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap),
                          ctypes.c_int(gap_ext), out.ctypes.data_as(ctypes.c_void_p),
                          cig.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(cap))
    if rc_:
        raise MemoryError("swo_align")
    d = {k: int(out[0][k]) for k in RESULT_DT.names}
    n = d.pop("n_cigar")
    d["cigar"] = [(int(c >> 2), _OPS[int(c & 3)]) for c in cig[:n]]
    return d


def find_all(read, seqs, thr, match=2, mismatch=-1, gap=-1, gap_ext=-1, cap=None):
    """Hits of one read against seqs (list of str) in the reference's pre-sort order.
    -> (hits structured array, stats dict)."""
    r = _buf(read.encode() if isinstance(read, str) else read)
    concat = np.frombuffer("".join(seqs).encode(), np.uint8)
    offs = np.zeros(len(seqs) + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    if cap is None:
        cap = max(64, 4 * len(r))
    while True:
        hits = np.zeros(cap, HIT_DT)
        st = np.zeros(1, STATS_DT)
        n = ctypes.c_int64(0)
        rc_ = lib().swo_find_all(r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(r)),
                                 ctypes.c_int(len(seqs)), concat.ctypes.data_as(ctypes.c_void_p),
                                 offs.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap),
                                 ctypes.c_int(gap_ext), ctypes.c_double(thr),
                                 hits.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap),
                                 ctypes.byref(n), st.ctypes.data_as(ctypes.c_void_p))
        if rc_:
            raise MemoryError("swo_find_all")
        if n.value <= cap:
            return hits[:n.value], {"calls": int(st[0]["calls"]), "cells": int(st[0]["cells"]),
                                    "max_depth": int(st[0]["max_depth"])}
        cap = n.value
