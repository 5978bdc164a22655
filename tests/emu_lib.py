"""TEST INFRASTRUCTURE: ctypes binding of liblsw_emu.so, the host emulation of the CUDA kernels'
lane logic (lamprey_b200/csrc/lsw_emu.cu).  Lets the CPU test tier check the kernel arithmetic
against the oracle.  Never used by the product path."""
import ctypes
import os

import numpy as np

from lamprey_b200._native import HIT_DT, RESULT_DT, TASK_DT, concat_sequences, _ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.environ.get("LSW_EMU_LIB") or os.path.join(ROOT, "lamprey_b200", "liblsw_emu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(PATH)
        _lib.lsw_emu_find_all.restype = ctypes.c_int
        _lib.lsw_emu_align_tasks.restype = ctypes.c_int
    return _lib


def find_all(reads, seqs, thr, match=2, mismatch=-1, gap=-1, force_full=False, allowed_overlap=-1, reuse=True, cap=None):
    rc_, ro = concat_sequences(reads)
    qc, qo = concat_sequences(seqs)
    qo = qo.astype(np.int32)
    cap = max(1024, int(ro[-1]) // 4) if cap is None else cap
    out = np.zeros(cap, HIT_DT)
    n = ctypes.c_int64(0)
    cnt = np.zeros(8, np.int64)
    rc = lib().lsw_emu_find_all(ctypes.c_int64(len(reads)), _ptr(rc_), _ptr(ro), ctypes.c_int(len(seqs)), _ptr(qc),
                                _ptr(qo), ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap),
                                ctypes.c_double(thr), _ptr(out), ctypes.c_int64(cap), ctypes.byref(n), _ptr(cnt), ctypes.c_int(int(force_full)), ctypes.c_int(int(allowed_overlap)), ctypes.c_int(int(reuse)))
    if rc == -4 and n.value > cap:       # LSW_E_OVERFLOW: short sequences on low-complexity reads give more than a hit per 4 bases
        return find_all(reads, seqs, thr, match, mismatch, gap, force_full, allowed_overlap, reuse, cap=n.value + 16)
    if rc:
        raise RuntimeError("lsw_emu_find_all rc=%d" % rc)
    return out[:n.value], dict(tasks=int(cnt[0]), cells=int(cnt[1]), steps=int(cnt[2]), rounds=int(cnt[3]), traced=int(cnt[4]), fallbacks=int(cnt[5]), computed_cells=int(cnt[6]), tier2=int(cnt[7]))


def align_tasks(reads, seqs, tasks, match=2, mismatch=-1, gap=-1, force_full=False, stats=None):
    rc_, ro = concat_sequences(reads)
    qc, qo = concat_sequences(seqs)
    qo = qo.astype(np.int32)
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DT)
    out = np.zeros(len(tasks), RESULT_DT)
    cap = max(64, 8 * len(tasks) + int(sum(t["end"] - t["start"] for t in tasks)))
    cig = np.zeros(cap, np.uint32)
    used = ctypes.c_int64(0)
    fb = ctypes.c_int64(0)
    rc = lib().lsw_emu_align_tasks(ctypes.c_int64(len(reads)), _ptr(rc_), _ptr(ro), ctypes.c_int(len(seqs)), _ptr(qc),
                                   _ptr(qo), ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap),
                                   ctypes.c_int64(len(tasks)), _ptr(tasks), _ptr(out), _ptr(cig), ctypes.c_int64(cap),
                                   ctypes.byref(used), ctypes.c_int(int(force_full)), ctypes.byref(fb))
    if rc:
        raise RuntimeError("lsw_emu_align_tasks rc=%d" % rc)
    if stats is not None:
        stats["fallbacks"] = fb.value
    return out, cig[:used.value]


def ssw_find_all(reads, seqs, thr, match=2, mismatch=-3, gap_open=5, gap_extend=2, packed=True, cap=None):
    """The default (skbio / SSW) seeding path through the host emulation of lsw_ssw.cuh."""
    rc_, ro = concat_sequences(reads)
    qc, qo = concat_sequences(seqs)
    qo = qo.astype(np.int32)
    cap = max(1024, int(ro[-1]) // 4) if cap is None else cap
    out = np.zeros(cap, HIT_DT)
    n = ctypes.c_int64(0)
    cnt = np.zeros(4, np.int64)
    lib().lsw_emu_ssw_find_all.restype = ctypes.c_int
    rc = lib().lsw_emu_ssw_find_all(ctypes.c_int64(len(reads)), _ptr(rc_), _ptr(ro), ctypes.c_int(len(seqs)), _ptr(qc), _ptr(qo),
                                    ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap_open), ctypes.c_int(gap_extend),
                                    ctypes.c_double(thr), _ptr(out), ctypes.c_int64(cap), ctypes.byref(n), _ptr(cnt), ctypes.c_int(int(packed)))
    if rc == -4 and n.value > cap:
        return ssw_find_all(reads, seqs, thr, match, mismatch, gap_open, gap_extend, packed, cap=n.value + 16)
    if rc:
        raise RuntimeError("lsw_emu_ssw_find_all rc=%d" % rc)
    return out[:n.value], dict(tasks=int(cnt[0]), cells=int(cnt[1]), rounds=int(cnt[2]), tier2=int(cnt[3]))


def ssw_align_tasks(reads, seqs, tasks, match=2, mismatch=-3, gap_open=5, gap_extend=2, packed=True):
    rc_, ro = concat_sequences(reads)
    qc, qo = concat_sequences(seqs)
    qo = qo.astype(np.int32)
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DT)
    out = np.zeros(len(tasks), RESULT_DT)
    cap = max(64, 8 * len(tasks) + int(sum(t["end"] - t["start"] for t in tasks)))
    cig = np.zeros(cap, np.uint32)
    used = ctypes.c_int64(0)
    lib().lsw_emu_ssw_align_tasks.restype = ctypes.c_int
    rc = lib().lsw_emu_ssw_align_tasks(ctypes.c_int64(len(reads)), _ptr(rc_), _ptr(ro), ctypes.c_int(len(seqs)), _ptr(qc), _ptr(qo),
                                       ctypes.c_int(match), ctypes.c_int(mismatch), ctypes.c_int(gap_open), ctypes.c_int(gap_extend),
                                       ctypes.c_int64(len(tasks)), _ptr(tasks), _ptr(out), _ptr(cig), ctypes.c_int64(cap), ctypes.byref(used),
                                       ctypes.c_int(int(packed)))
    if rc:
        raise RuntimeError("lsw_emu_ssw_align_tasks rc=%d" % rc)
    return out, cig[:used.value]


def ssw_subreads(reads, pairs, target, match=2, mismatch=-3, gap_open=5, gap_extend=2):
    """Stage 2 (query = sub-read of any length, target = reference) through the host emulation of ssw_generic_task."""
    from lamprey_b200._native import PAIR_DT
    rc_, ro = concat_sequences(reads)
    pairs = np.ascontiguousarray(pairs, dtype=PAIR_DT)
    t = np.frombuffer(target.encode(), np.uint8)
    out = np.zeros(len(pairs), RESULT_DT)
    cap = max(64, 16 * len(pairs) + 4 * int(sum(p["end"] - p["start"] for p in pairs)))
    cig = np.zeros(cap, np.uint32)
    used = ctypes.c_int64(0)
    lib().lsw_emu_ssw_subreads.restype = ctypes.c_int
    rc = lib().lsw_emu_ssw_subreads(ctypes.c_int64(len(reads)), _ptr(rc_), _ptr(ro), ctypes.c_int(match), ctypes.c_int(mismatch),
                                    ctypes.c_int(gap_open), ctypes.c_int(gap_extend), ctypes.c_int64(len(pairs)), _ptr(pairs), _ptr(t),
                                    ctypes.c_int(len(t)), _ptr(out), _ptr(cig), ctypes.c_int64(cap), ctypes.byref(used))
    if rc:
        raise RuntimeError("lsw_emu_ssw_subreads rc=%d" % rc)
    return out, cig[:used.value]
