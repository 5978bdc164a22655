"""FASTQ -> (uint8 bases, int64 offsets) without Biopython.

LAMPrey reads its input with ``SeqIO.parse(fn, "fastq")`` (/root/reference/lamprey.py:2145-2155) and hands
``str(record.seq)`` to every align() call.  Once seeding runs at ~10^6 reads/s that parser is the bottleneck,
so this reader slices the sequence lines straight out of the file bytes with numpy: the result is exactly
the ``(ascii, offsets)`` pair ``lsw_upload_reads`` takes, with no per-read Python objects.
Four-line records only (the format ONT basecallers and LAMPrey's fixtures use).
"""
import gzip

import numpy as np


def read_fastq(path, with_names=False):
    """-> (concat uint8, offsets int64[n+1]) [, names list].  The FASTQ must use 4-line records."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fp:
        raw = np.frombuffer(fp.read(), dtype=np.uint8)
    if raw.size == 0:
        return (np.zeros(0, np.uint8), np.zeros(1, np.int64)) + (([],) if with_names else ())
    nl = np.flatnonzero(raw == 10)
    if raw[-1] != 10:
        nl = np.append(nl, raw.size)
    if len(nl) % 4:
        raise ValueError("%s: %d lines is not a multiple of 4 (multi-line FASTQ records are not supported)" % (path, len(nl)))
    starts = np.concatenate(([0], nl[:-1] + 1))
    if not (raw[starts[0::4]] == ord("@")).all() or not (raw[starts[2::4]] == ord("+")).all():
        raise ValueError("%s: not a 4-line FASTQ" % path)
    s, e = starts[1::4], nl[1::4].copy()
    e -= (raw[np.maximum(e - 1, 0)] == 13) & (e > s)          # tolerate CRLF
    lens = (e - s).astype(np.int64)
    offs = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    # gather the sequence bytes: index = start of own line + position inside the read
    idx = np.repeat(s - offs[:-1], lens) + np.arange(offs[-1], dtype=np.int64)
    concat = raw[idx]
    if not with_names:
        return concat, offs
    hs, he = starts[0::4], nl[0::4]
    names = [raw[a + 1:b].tobytes().decode("ascii", "replace").split()[0] if b > a + 1 else "" for a, b in zip(hs, he)]
    return concat, offs, names
