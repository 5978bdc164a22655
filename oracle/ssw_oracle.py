"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes loader for oracle/libssw_oracle.so (ssw_oracle.c: the restatement of
scikit-bio's StripedSmithWaterman / SSW behind /root/reference/lamprey.py:572-593 and 1651-1658), plus the thin Python
layer of skbio's wrapper that LAMPrey touches (AlignmentStructure fields, aligned sequences, CIGAR string) and the
restated alignSequence_skbio.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
PARITY UNPINNED: scikit-bio is absent here; see the header of ssw_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "libssw_oracle.so")
_lib = None

RESULT_DT = np.dtype([("score", "<i4"), ("ref_begin", "<i4"), ("ref_end", "<i4"), ("read_begin", "<i4"), ("read_end", "<i4"),
                      ("score2", "<i4"), ("ref_end2", "<i4"), ("n_cigar", "<i4"), ("matches", "<i4")])
HIT_DT = np.dtype([("seq", "<i4"), ("start", "<i4"), ("end", "<i4"), ("matches", "<i4"), ("score", "<i4"), ("q_pos", "<i4"),
                   ("q_end", "<i4"), ("depth", "<i4")])
STATS_DT = np.dtype([("calls", "<i8"), ("cells", "<i8"), ("max_depth", "<i4"), ("_pad", "<i4")])
_OPS = " MID"


def build(force=False):
    src = os.path.join(HERE, "ssw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.sswo_align.restype = ctypes.c_int
        _lib.sswo_find_all.restype = ctypes.c_int
    return _lib


def _buf(s):
    return np.frombuffer(s.encode() if isinstance(s, str) else bytes(s), dtype=np.uint8)


def align(target, query, match=2, mismatch=-3, gap_open=5, gap_extend=2, mask_length=None):
    """ssw_align(query profile, target) with skbio's default flag -> dict of the raw SSW fields + cigar [(n, op)]."""
    t, q = _buf(target), _buf(query)
    out = np.zeros(1, RESULT_DT)
    cap = len(t) + len(q) + 8
    cig = np.zeros(cap, np.uint32)
    if mask_length is None:
        mask_length = max(15, len(q) // 2)
    rc = lib().sswo_align(t.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(t)), q.ctypes.data_as(ctypes.c_void_p),
                          ctypes.c_int(len(q)), ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap_open),
                          ctypes.c_int(gap_extend), ctypes.c_int(mask_length), out.ctypes.data_as(ctypes.c_void_p),
                          cig.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(cap))
    if rc == -1:
        raise MemoryError("sswo_align")
    d = {k: int(out[0][k]) for k in RESULT_DT.names}
    d["banded_failed"] = rc == -2
    n = d.pop("n_cigar")
    d["cigar"] = [(int(c >> 2), _OPS[int(c & 3)]) for c in cig[:n]]
    return d


def find_all(read, seqs, thr, match=2, mismatch=-3, gap_open=5, gap_extend=2, cap=None):
    """Hits of one read (default seeding path, lamprey.py:596-645 with args.swalign False) in pre-sort order."""
    r = _buf(read)
    concat = np.frombuffer("".join(seqs).encode(), np.uint8)
    offs = np.zeros(len(seqs) + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    if cap is None:
        cap = max(64, 4 * len(r))
    while True:
        hits = np.zeros(cap, HIT_DT)
        st = np.zeros(1, STATS_DT)
        n = ctypes.c_int64(0)
        rc = lib().sswo_find_all(r.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(r)), ctypes.c_int(len(seqs)),
                                 concat.ctypes.data_as(ctypes.c_void_p), offs.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap_open), ctypes.c_int(gap_extend),
                                 ctypes.c_double(thr), hits.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap),
                                 ctypes.byref(n), st.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise MemoryError("sswo_find_all")
        if n.value <= cap:
            return hits[:n.value], {"calls": int(st[0]["calls"]), "cells": int(st[0]["cells"]), "max_depth": int(st[0]["max_depth"])}
        cap = n.value


class AlignmentStructure(object):
    """The fields of skbio.alignment.AlignmentStructure that LAMPrey reads (zero_index=True: 0-based, inclusive ends)."""

    def __init__(self, d, query, target):
        self.optimal_alignment_score = d["score"]
        self.suboptimal_alignment_score = d["score2"]
        self.target_begin = d["ref_begin"]
        self.target_end_optimal = d["ref_end"]
        self.target_end_suboptimal = d["ref_end2"]
        self.query_begin = d["read_begin"]
        self.query_end = d["read_end"]
        self.query_sequence = query
        self.target_sequence = target
        self._cigar = d["cigar"]
        self.cigar = "".join("%d%s" % x for x in d["cigar"])

    def _aligned(self, seq, begin, end, gap_type):
        # skbio _get_aligned_sequence: 'M' and the other gap type consume the sequence, its own gap type writes '-';
        # whatever of [begin, end] the CIGAR did not consume is appended
        s = seq[begin:end + 1]
        out, index = [], 0
        for length, op in self._cigar:
            if op == gap_type:
                out.append('-' * length)
            else:
                out.append(s[index:index + length])
                index += length
        out.append(s[index:end - begin + 1])
        return "".join(out)

    @property
    def aligned_query_sequence(self):
        return self._aligned(self.query_sequence, self.query_begin, self.query_end, 'D')

    @property
    def aligned_target_sequence(self):
        return self._aligned(self.target_sequence, self.target_begin, self.target_end_optimal, 'I')


class StripedSmithWaterman(object):
    """skbio.alignment.StripedSmithWaterman(query)(target) with its default parameters."""

    def __init__(self, query_sequence, gap_open_penalty=5, gap_extend_penalty=2, match_score=2, mismatch_score=-3):
        self.query = str(query_sequence)
        self.p = (match_score, mismatch_score, gap_open_penalty, gap_extend_penalty)

    def __call__(self, target_sequence):
        target = str(target_sequence)
        return AlignmentStructure(align(target, self.query, *self.p), self.query, target)


def align_sequence_skbio(seq, primer, primer_name):
    """/root/reference/lamprey.py:572-593 -> (start, end, identity, name)."""
    a = StripedSmithWaterman(primer)(str(seq))
    aq, at = a.aligned_query_sequence, a.aligned_target_sequence
    matches = 0
    for i in range(len(aq)):
        if aq[i] == at[i]:
            matches += 1
    return (a.target_begin, a.target_end_optimal + 1, float(matches) / float(len(primer)), primer_name)
