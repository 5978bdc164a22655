"""Next rows of SURVEY.md section 8f on the host: the chain step that consumes the seed hits and the FASTQ reader."""
import gzip
import os
import random

import numpy as np
import pytest

from conftest import load_reads, schema_path, load_seed_golden
from lamprey_b200 import chain, fastq, seeding
from oracle import chain_ref, seeding_ref


def _golden_alignments(cls, reads_key="50", schema="H33", thr=0.70):
    gold = load_seed_golden(reads_key, schema, thr)
    names = [str(n) for n in gold["names"]]
    g = gold["hits"]
    out = []
    for i in range(int(g[:, 0].max()) + 1):
        rows = g[g[:, 0] == i]
        out.append([cls(int(r[2]), int(r[3]), r[4] / (r[4] + r[5]), names[r[1]]) for r in rows])
    return out


def test_remove_overlaps_and_amplicon_match_oracle():
    ours = _golden_alignments(seeding.Alignment)
    ref = _golden_alignments(seeding_ref.Alignment)
    ps_o, ps_r = seeding.PrimerSet(schema_path("H33")), seeding_ref.PrimerSet(schema_path("H33"))
    reads = load_reads("50")
    n_targets = 0
    for i, (a, b) in enumerate(zip(ours, ref)):
        if not a:
            continue
        pa = chain.removeOverlappingPrimerAlignments(False, a, 4)          # lamprey.py:1536
        pb = chain_ref.remove_overlapping_primer_alignments(b, 4)
        assert [(x.start, x.end, x.primer_name) for x in pa] == [(x.start, x.end, x.primer_name) for x in pb]
        for x, y in zip(pa, pb):
            if x.primer_name in ("T", "Tc"):
                n_targets += 1
                assert chain.extractAmpliconAroundTarget(ps_o, pa, x) == chain_ref.extract_amplicon_around_target(ps_r, pb, y)
        subs = chain.target_subreads(ps_o, pa, len(reads[i]), i)
        assert len(subs) == sum(1 for x in pa if x.primer_name in ("T", "Tc"))
    assert n_targets > 20


def test_remove_overlaps_random():
    rng = random.Random(4)
    for _ in range(200):
        n = rng.randrange(1, 30)
        starts = sorted(rng.randrange(0, 300) for _ in range(n))
        hits = [(s, s + rng.randrange(10, 40), rng.choice([0.7, 0.75, 0.8, 0.9, 1.0]), rng.choice("ABCD")) for s in starts]
        a = chain.removeOverlappingPrimerAlignments(False, [seeding.Alignment(*h) for h in hits], 4)
        b = chain_ref.remove_overlapping_primer_alignments([seeding_ref.Alignment(*h) for h in hits], 4)
        assert [(x.start, x.end, x.identity, x.primer_name) for x in a] == [y.astuple() for y in b]


def test_read_fastq(tmp_path):
    reads = load_reads("50")
    p = tmp_path / "x.fastq"
    with open(p, "w") as fp:
        for i, r in enumerate(reads):
            fp.write("@read%d runid=abc ch=%d start_time=2021-02-08T02:45:38Z\n%s\n+\n%s\n" % (i, i, r, "#" * len(r)))
    concat, offs, names = fastq.read_fastq(str(p), with_names=True)
    c2, o2 = fastq.read_fastq(str(p), mmap=False, threads=2)
    assert (c2 == concat).all() and (o2 == offs).all()
    assert names[:2] == ["read0", "read1"] and len(offs) == len(reads) + 1
    got = [concat[offs[i]:offs[i + 1]].tobytes().decode() for i in range(len(reads))]
    assert got == reads
    pz = tmp_path / "y.fastq.gz"
    with gzip.open(pz, "wt") as fp:
        fp.write("@a\nACGT\r\n+\nIIII\r\n@b\n\n+\n\n@c\nTTGCA\n+\nIIIII")          # CRLF, empty read, no final newline
    concat, offs = fastq.read_fastq(str(pz))
    assert [concat[offs[i]:offs[i + 1]].tobytes() for i in range(3)] == [b"ACGT", b"", b"TTGCA"]


def test_emu_overlap_pruning_matches_reference_function():
    import emu_lib
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    names = [n for n, _ in ps.sequences()]
    hits, _ = emu_lib.find_all(reads, [s for _, s in ps.sequences()], 0.70, allowed_overlap=4)
    for i in range(len(reads)):
        h = hits[hits["read"] == i]
        if not len(h):
            continue
        al = [seeding.Alignment(int(x["start"]), int(x["end"]), x["matches"] / (x["matches"] + x["mismatches"]), names[x["query"]]) for x in h]
        want = chain.removeOverlappingPrimerAlignments(False, al, 4)
        got = [a for a, x in zip(al, h) if x["reserved"] == 1]
        assert [(a.start, a.end, a.primer_name) for a in got] == [(a.start, a.end, a.primer_name) for a in want]


def test_emu_device_subreads_match_reference_functions():
    # lsw_chain.cuh (the device's extractAmpliconAroundTarget / stage-1 loop) on the host against the host mirror and the oracle
    import ctypes
    import emu_lib
    from lamprey_b200 import _native
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    ps_r = seeding_ref.PrimerSet(schema_path("H33"))
    names = [n for n, _ in ps.sequences()]
    index = {n: k for k, n in enumerate(names)}
    hits, _ = emu_lib.find_all(reads, [s for _, s in ps.sequences()], 0.70, allowed_overlap=4)
    hits = np.ascontiguousarray(hits)
    fo = np.array([index.get(n, -1) for n in ps.fwd_primer_order], np.int32)
    ro = np.array([index.get(n, -1) for n in ps.rev_primer_order], np.int32)
    rlen = np.array([len(r) for r in reads], np.int32)
    out = np.zeros(4096, _native.SUBREAD_DT)
    n = ctypes.c_int64(0)
    lib = emu_lib.lib()
    lib.lsw_emu_chain_subreads.restype = ctypes.c_int
    rc = lib.lsw_emu_chain_subreads(emu_lib._ptr(hits), ctypes.c_int64(len(hits)), ctypes.c_int64(len(reads)), emu_lib._ptr(rlen),
                                    ctypes.c_int(len(fo)), emu_lib._ptr(fo), ctypes.c_int(len(ro)), emu_lib._ptr(ro),
                                    ctypes.c_int(index["T"]), ctypes.c_int(index["Tc"]), emu_lib._ptr(out), ctypes.c_int64(len(out)),
                                    ctypes.byref(n))
    assert rc == 0 and n.value > 20
    got = out[:n.value]
    k = 0
    for i in range(len(reads)):
        h = hits[(hits["read"] == i) & (hits["reserved"] == 1)]
        al = [seeding.Alignment(int(x["start"]), int(x["end"]), x["matches"] / (x["matches"] + x["mismatches"]), names[x["query"]]) for x in h]
        want = chain.target_subreads(ps, al, len(reads[i]), i)
        al_r = [seeding_ref.Alignment(a.start, a.end, a.identity, a.primer_name) for a in al]
        for a, w in zip([a for a in al_r if a.primer_name in ("T", "Tc")], want):
            s, e = chain_ref.extract_amplicon_around_target(ps_r, al_r, a)
            assert (s, len(reads[i]) - 1 if e == -1 else e) == w[:2]
        mine = [(int(x["start"]), int(x["end"]), "%d_%d" % (i, x["target_idx"]), "fwd" if x["direction"] == 0 else "rev")
                for x in got[k:k + len(want)]]
        assert mine == want and all(x["read"] == i for x in got[k:k + len(want)])
        k += len(want)
    assert k == n.value


# ---- host side of the path: FASTQ scanner, hit-list index, record expansion (lsw_hostops.cpp), against plain numpy ----
def _fastq_numpy(raw):
    """Independent restatement of the scanner for the tests: newline positions with numpy."""
    raw = np.frombuffer(raw, np.uint8)
    while raw.size and raw[-1] in (10, 13):
        raw = raw[:-1]
    if raw.size == 0:
        return [], []
    nl = np.append(np.flatnonzero(raw == 10), raw.size)
    starts = np.concatenate(([0], nl[:-1] + 1))
    lines = [raw[a:b].tobytes().rstrip(b"\r") for a, b in zip(starts, nl)]
    if len(lines) % 4 == 3 and lines[-1].startswith(b"+"):          # the empty quality line of a last, empty read was stripped
        lines.append(b"")
    assert len(lines) % 4 == 0
    return [l[1:].split()[0].decode() if len(l) > 1 else "" for l in lines[0::4]], lines[1::4]


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_fastq_scanner_matches_numpy_on_ragged_files(threads, tmp_path):
    rng = np.random.default_rng(threads)
    for case in range(6):
        n = [0, 1, 2, 57, 400, 3000][case]
        recs = []
        for i in range(n):
            L = int(rng.choice([0, 1, 5, 40, 400, 3000]))
            seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), L))
            qual = bytes(rng.choice(np.frombuffer(b"@+#I5", np.uint8), L))           # '@' and '+' inside quality strings
            eol = b"\r\n" if case == 3 else b"\n"
            recs.append(b"@r%d extra=%d" % (i, i) + eol + seq + eol + b"+" + eol + qual + eol)
        data = b"".join(recs)
        if case in (2, 4):
            data = data.rstrip(b"\r\n")                                              # no final newline
        if case == 5:
            data += b"\n\n"                                                          # blank lines at the end
        want_names, want_seqs = _fastq_numpy(data)
        concat, offs, names = fastq.parse_fastq(data, with_names=True, threads=threads)
        assert len(offs) == n + 1 and offs[0] == 0 and len(names) == n
        assert [concat[offs[i]:offs[i + 1]].tobytes() for i in range(n)] == want_seqs
        assert list(names) == want_names
    # a big single-line-heavy file cut into many pieces: every piece boundary falls inside a line
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(L))) for L in rng.integers(1, 200000, 40)]
    data = b"".join(b"@x\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for s in seqs)
    concat, offs = fastq.parse_fastq(data, threads=threads)
    assert [concat[offs[i]:offs[i + 1]].tobytes() for i in range(len(seqs))] == seqs


def test_fastq_scanner_refuses_what_biopython_refuses(tmp_path):
    good = b"@a\nACGT\n+\nIIII\n"
    for bad, what in ((b"@a\nACGT\n+\nIII\n", "differ"), (b"a\nACGT\n+\nIIII\n", "'@'"), (b"@a\nACGT\n-\nIIII\n", "third line"),
                      (good + b"@b\nAC\n+\n", "differ"), (good + b"@b\nAC\n", "multiple of 4"), (good + b"@b\nAC\nGT\n+\nIIII\n", "multiple of 4"),
                      (good + b"\n" + good, "multiple of 4")):
        with pytest.raises(ValueError, match=what):
            fastq.parse_fastq(good * 3 + bad)
    p = tmp_path / "bad.fastq"
    p.write_bytes(b"@a\nACGT\n+\nII\n")
    with pytest.raises(ValueError, match="bad.fastq.*record 1"):
        fastq.read_fastq(str(p))
    p = tmp_path / "empty.fastq"
    p.write_bytes(b"")
    concat, offs = fastq.read_fastq(str(p))
    assert len(concat) == 0 and list(offs) == [0]


def test_hit_list_index_and_expansion_match_numpy():
    from lamprey_b200 import _native
    rng = np.random.default_rng(11)
    for n_reads, per in ((0, 0), (1, 0), (5, 3), (3000, 40)):
        counts = rng.integers(0, per + 1, n_reads)                      # reads without hits in between
        n = int(counts.sum())
        c = np.zeros(n, _native.HIT16_DT)
        c["read"] = np.repeat(np.arange(n_reads), counts)
        c["start"] = rng.integers(0, 1 << 20, n); c["span"] = rng.integers(0, 190, n); c["query"] = rng.integers(0, 46, n)
        c["depth"] = rng.integers(0, 65536, n); c["matches"] = rng.integers(0, 256, n); c["mismatches"] = rng.integers(0, 64, n)
        c["q_pos"] = rng.integers(0, 20, n); c["q_end"] = c["q_pos"] + rng.integers(0, 43, n)
        for threads in (1, 5):
            w = _native.expand_hits(c, 2, -1, -1, threads=threads)
            m = (c["matches"] & 0x7F).astype(np.int64); x = c["mismatches"].astype(np.int64); span = c["span"].astype(np.int64)
            qspan = c["q_end"].astype(np.int64) - c["q_pos"]
            a = span + qspan - m - x
            assert (w["read"] == c["read"]).all() and (w["start"] == c["start"]).all() and (w["end"] == c["start"].astype(np.int64) + span).all()
            assert (w["query"] == c["query"]).all() and (w["depth"] == np.minimum(c["depth"], 32767)).all()
            assert (w["matches"] == m).all() and (w["mismatches"] == x).all() and (w["reserved"] == (c["matches"] >> 7)).all()
            assert (w["score"] == (2 * m - (a - m) - ((span - a) + (qspan - a))).astype(np.int16)).all()
            assert (w["q_pos"] == c["q_pos"]).all() and (w["q_end"] == c["q_end"]).all() and (w["reserved2"] == 0).all()
            for hits in (c, w):
                offs, cov = _native.hits_index(hits, n_reads, threads=threads)
                assert (offs == np.searchsorted(c["read"], np.arange(n_reads + 1))).all()
                want = np.zeros(n_reads, np.int64)
                np.add.at(want, c["read"].astype(np.int64), span)
                assert (cov == want).all()
    # read ids with a base (a range of a split batch), unsorted lists and ids outside the batch are refused
    c = np.zeros(4, _native.HIT16_DT)
    c["read"] = [10, 10, 12, 13]; c["span"] = 7
    offs, cov = _native.hits_index(c, 5, read_base=10)
    assert list(offs) == [0, 2, 2, 3, 4, 4] and list(cov) == [14, 0, 7, 7, 0]
    with pytest.raises(ValueError):
        _native.hits_index(c, 3, read_base=10)
    c["read"] = [3, 2, 2, 1]
    with pytest.raises(ValueError):
        _native.hits_index(c, 5)
