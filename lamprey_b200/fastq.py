"""FASTQ -> (uint8 bases, int64 offsets) without Biopython.

LAMPrey reads its input with ``SeqIO.parse(fn, "fastq")`` (/root/reference/lamprey.py:2145-2155; one delivered file at a
time in online mode, lamprey.py:451-507) and hands ``str(record.seq)`` to every align() call.  With seeding at ~2 * 10^6
reads/s per GPU that parser -- and any per-read Python object -- is the bottleneck, so the file's bytes go through the
library's multi-threaded scanner (``lsw_fastq_index`` / ``lsw_fastq_extract``, lamprey_b200/csrc/lsw_hostops.cpp): the
result is exactly the ``(ascii, offsets)`` pair ``lsw_upload_reads`` / ``lsw_pool_submit`` take.
Four-line records only (the format ONT basecallers and LAMPrey's fixtures use); what Biopython raises ``ValueError`` for
(no '@' / '+', sequence and quality of different length, a truncated record) raises ``ValueError`` here too.
"""
import ctypes
import gzip

import numpy as np

from . import _native


class ReadNames(object):
    """Read ids of a parsed FASTQ (Biopython's ``record.id``: the title up to the first white space), decoded on demand."""

    def __init__(self, data, spans):
        self._data, self._spans = data, spans

    def __len__(self):
        return len(self._spans)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        a, b = self._spans[i]
        return self._data[a:b].tobytes().decode("ascii", "replace")

    def __iter__(self):
        return (self[k] for k in range(len(self)))


def parse_fastq(data, with_names=False, threads=0, pinned=False):
    """data: the bytes of a FASTQ file (bytes / uint8 array / memory map).  -> (concat uint8, offsets int64[n+1]) [, names].
    pinned=True puts both arrays into page-locked memory (``_native.PinnedArray``: H2D at PCIe speed; needs a CUDA device)
    and returns the two PinnedArray objects instead of plain arrays."""
    lib = _native.load()
    raw = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    raw = np.ascontiguousarray(raw, np.uint8)
    h = ctypes.c_void_p(0)
    n, nb = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.lsw_fastq_index(_native._ptr(raw), raw.size, int(threads), ctypes.byref(h), ctypes.byref(n), ctypes.byref(nb))
    if rc != _native.LSW_OK:
        raise ValueError(lib.lsw_host_last_error().decode())
    try:
        if pinned:
            concat_p, offs_p = _native.PinnedArray(nb.value, np.uint8), _native.PinnedArray(n.value + 1, np.int64)
            concat, offs = concat_p.array, offs_p.array
        else:
            concat, offs = np.empty(nb.value, np.uint8), np.empty(n.value + 1, np.int64)
        spans = np.empty((n.value, 2), np.int64) if with_names else None
        rc = lib.lsw_fastq_extract(h, _native._ptr(concat), ctypes.c_void_p(offs.ctypes.data), _native._ptr(spans) if with_names else None)
        if rc != _native.LSW_OK:
            raise ValueError(lib.lsw_host_last_error().decode())
    finally:
        lib.lsw_fastq_free(h)
    out = (concat_p, offs_p) if pinned else (concat, offs)
    return out + ((ReadNames(raw, spans),) if with_names else ())


def read_fastq(path, with_names=False, threads=0, pinned=False, mmap=True):
    """-> (concat uint8, offsets int64[n+1]) [, names].  The FASTQ must use 4-line records; ``.gz`` files are inflated first.
    mmap=False reads the file into memory first instead of mapping it (slower by the copy; use it for a file another process may
    still truncate or rewrite -- a mapped page that disappears is a bus error, not an exception)."""
    import os
    if str(path).endswith(".gz"):
        with gzip.open(path, "rb") as fp:
            data = fp.read()
    elif not os.path.getsize(path):
        data = np.zeros(0, np.uint8)
    elif mmap:
        # mapped, not read: the scanner's threads pull the pages straight from the page cache (np.fromfile costs as much as the parse)
        data = np.memmap(path, dtype=np.uint8, mode="r")
    else:
        data = np.fromfile(path, dtype=np.uint8)
    try:
        return parse_fastq(data, with_names=with_names, threads=threads, pinned=pinned)
    except ValueError as e:
        raise ValueError("%s: %s" % (path, e)) from None
