"""Next rows of SURVEY.md section 8f on the host: the chain step that consumes the seed hits and the FASTQ reader."""
import gzip
import os
import random

import numpy as np

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
