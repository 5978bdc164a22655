"""Host-side mirror of the reference interface (no GPU): schema parsing, rc, ordering, sharding,
pre-order reconstruction, and the 2-rank gloo gather used by bench.py."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT, schema_path, load_seed_golden
from lamprey_b200 import seeding, synth
from lamprey_b200._native import HIT_DT
from oracle import seeding_ref, swalign_ref


def test_rc_and_primerset_match_oracle():
    for name in ("H33", "H31"):
        a, b = seeding.PrimerSet(schema_path(name)), seeding_ref.PrimerSet(schema_path(name))
        assert a.primer_dict == b.primer_dict
        assert list(a.primer_dict) == list(b.primer_dict)
        assert (a.fwd_primer_order, a.rev_primer_order) == (b.fwd_primer_order, b.rev_primer_order)
        assert a.sequences() == b.sequences()
    assert seeding.rc("ACGTNacgt") == seeding_ref.rc("ACGTNacgt") == "tgcaNACGT"
    # names ending in c are stored reverse-complemented under the stripped name (lamprey.py:129-130)
    p = seeding.PrimerSet(schema_path("H31"))
    assert "B3" in p.primer_dict and "B3c" not in p.primer_dict
    assert len(p.sequences()) == 46 and p.sequences()[0][0] == "T" and p.sequences()[1][0] == "Tc"


def test_alignment_record_semantics():
    a, b = seeding.Alignment(3, 9, 0.5, "T"), seeding.Alignment(5, 7, 0.9, "F3")
    assert a < b and not (b < a)
    assert (a == b) is None                      # the reference's __eq__ returns None
    assert str(a) == "T at (3, 9) identity: 0.5"
    assert [x.primer_name for x in sorted([b, a])] == ["T", "F3"]


def test_preorder_reconstruction_matches_reference_recursion():
    # pre-order list of findPrimerAlignments rebuilt from (start-sorted hits, depth)
    al = swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1))
    concat, offs = synth.generate(3, seed=11, mean_len=1500)
    for read in synth.reads_as_strings(concat, offs):
        primer = "GCTCGCAAGAGTGCG"
        want = seeding_ref.find_primer_alignments(al, read, primer, "T", 0.7)

        def depths(seq, shift, d, out):
            a = al.align(seq, primer)
            al_len = a.r_end - a.r_pos
            if a.identity >= 0.7 and al_len >= len(primer) * 0.7:
                out.append((a.r_pos + shift, a.r_end + shift, a.matches, a.mismatches, d))
                if a.r_pos >= len(primer):
                    depths(seq[:a.r_pos], shift, d + 1, out)
                if len(seq) - a.r_end >= len(primer):
                    depths(seq[a.r_end:], shift + a.r_end, d + 1, out)
        rows = []
        depths(read, 0, 0, rows)
        rows.sort()
        hits = np.zeros(len(rows), HIT_DT)
        for k, (s, e, m, x, d) in enumerate(rows):
            hits[k]["start"], hits[k]["end"], hits[k]["matches"], hits[k]["mismatches"], hits[k]["depth"] = s, e, m, x, d
        ident = hits["matches"] / np.maximum(hits["matches"] + hits["mismatches"], 1)
        got = seeding._preorder(hits, ident, "T")
        assert [(h.start, h.end, h.identity) for h in got] == [(h.start, h.end, h.identity) for h in want]


def test_seedbatch_views_match_golden():
    gold = load_seed_golden("50", "H33", 0.75)
    g = gold["hits"]
    hits = np.zeros(len(g), HIT_DT)
    for k, col in enumerate(("read", "query", "start", "end", "matches", "mismatches", "score", "q_pos", "q_end", "depth")):
        hits[col] = g[:, k]
    names = [str(n) for n in gold["names"]]
    from conftest import load_reads
    lens = np.array([len(r) for r in load_reads("50")])
    batch = seeding.SeedBatch(hits, 50, lens, names, None)
    assert np.array_equal(batch.coverage, gold["coverage"])
    al, cov = batch[7]
    rows = g[g[:, 0] == 7]
    assert [(a.start, a.end, a.primer_name) for a in al] == [(r[2], r[3], names[r[1]]) for r in rows]
    assert cov == gold["coverage"][7]


def test_shard_reads_balances_bases():
    rng = np.random.default_rng(0)
    lens = rng.integers(50, 50000, 1000)
    offs = np.concatenate([[0], np.cumsum(lens)])
    for ws in (1, 2, 4, 8):
        sh = seeding.shard_reads(offs, ws)
        assert sh[0][0] == 0 and sh[-1][1] == 1000
        assert all(sh[i][1] == sh[i + 1][0] for i in range(ws - 1))
        per = [offs[hi] - offs[lo] for lo, hi in sh]
        assert max(per) - min(per) <= 2 * lens.max()


def test_two_rank_gloo_gather():
    # the multi-GPU path has no data-path collective: ranks seed disjoint read shards and rank 0
    # gathers hit counts / timings.  Exercised here with gloo, world_size 2, on CPU.
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, script], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]
