"""The oracle itself: restatement vs the known-answer vectors, C vs Python, Cython vs Python."""
import io
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLD, load_reads, schema_path, load_seed_golden
from oracle import swalign_ref, seeding_ref, c_oracle


def _al():
    return swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1))


def test_upstream_readme_vector():
    # the only known answer the reference carries (lamprey.py:1860-1865 copies the upstream README)
    a = _al().align('ACACACTA', 'AGCACACA')
    assert (a.score, a.cigar_str, a.matches, a.mismatches) == (12, '1M1I5M1D1M', 7, 2)
    assert (a.q_pos, a.q_end, a.r_pos, a.r_end) == (0, 8, 0, 8)
    assert abs(a.identity - 7 / 9) < 1e-15
    buf = io.StringIO()
    a.dump(out=buf)
    assert buf.getvalue() == ("Query: 1 AGCACAC-A 8\n         | ||||| |\nRef  : 1 A-CACACTA 8\n\n"
                              "Score: 12\nMatches: 7 (77.8%)\nMismatches: 2\nCIGAR: 1M1I5M1D1M\n")


def test_degenerate_cases():
    al = _al()
    a = al.align('AAAA', 'CCCC')
    assert (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end, a.cigar, a.identity) == (0, 4, 4, 4, 4, [], 0)
    assert isinstance(a.identity, int)
    a = al.align('', 'ACGT')
    assert (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end) == (0, 0, 0, 0, 0)
    a = al.align('ACGTACGTAC', 'ACGTACGTAC')
    assert (a.score, a.cigar_str) == (20, '10M')
    a = al.align('acgtNNacgt', 'ACGTNNACGT')            # upper-cased raw characters: N matches N
    assert (a.score, a.matches) == (20, 10)


def test_committed_vectors_python_and_c():
    vecs = json.load(open(os.path.join(GOLD, "align_vectors.json")))
    assert len(vecs) > 300
    al = _al()
    for v in vecs:
        a = al.align(v["ref"], v["query"])
        got = (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end, a.matches, a.mismatches, a.cigar_str)
        want = tuple(v[k] for k in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches", "cigar"))
        assert got == want
        c = c_oracle.align(v["ref"], v["query"])
        cg = (c["score"], c["q_pos"], c["q_end"], c["r_pos"], c["r_end"], c["matches"], c["mismatches"],
              "".join("%d%s" % x for x in c["cigar"]))
        assert cg == want


def test_c_vs_python_random_general_scoring():
    rng = random.Random(7)
    for it in range(400):
        match, mismatch, gap = rng.choice([(2, -1, -1), (1, -1, -1), (3, -2, -2), (2, 0, -1), (2, -3, -1)])
        al = swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(match, mismatch), gap, gap)
        alpha = rng.choice(["AC", "ACGT", "A", "ACGTN"])
        ref = "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 90)))
        q = "".join(rng.choice(alpha) for _ in range(rng.randrange(0, 25)))
        a = al.align(ref, q)
        c = c_oracle.align(ref, q, match, mismatch, gap, gap)
        assert (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end, a.matches, a.mismatches, a.cigar) == \
               (c["score"], c["q_pos"], c["q_end"], c["r_pos"], c["r_end"], c["matches"], c["mismatches"], c["cigar"])


def test_seeding_python_vs_golden_subset():
    # pure-Python oracle on a few shipped reads; the full sets are pinned through the C restatement below
    reads = load_reads("50")
    ps = seeding_ref.PrimerSet(schema_path("H33"))
    names = [n for n, _ in ps.sequences()]
    gold = load_seed_golden("50", "H33", 0.75)
    g = gold["hits"]
    for i in (0, 7, 23):
        hits, cov = seeding_ref.find_all_primer_alignments(_al(), reads[i], ps, 0.75)
        rows = g[g[:, 0] == i]
        assert [(names.index(h.primer_name), h.start, h.end) for h in hits] == [tuple(r[1:4]) for r in rows]
        assert [h.identity for h in hits] == [r[4] / (r[4] + r[5]) for r in rows]
        assert cov == gold["coverage"][i]


@pytest.mark.parametrize("schema", ["H33", "H31"])
@pytest.mark.parametrize("thr", [0.70, 0.75])
def test_seeding_c_vs_golden_50(schema, thr):
    reads = load_reads("50")
    ps = seeding_ref.PrimerSet(schema_path(schema))
    seqs = [s for _, s in ps.sequences()]
    gold = load_seed_golden("50", schema, thr)
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"],
                  x["q_end"], x["depth"]) for x in h]
    assert np.array_equal(np.array(rows, np.int32).reshape(-1, 10), gold["hits"])
    assert (calls, cells) == (int(gold["calls"]), int(gold["cells"]))


def test_survey_workload_numbers():
    # 3288 calls, 2.88e7 cells at thr 0.75
    g = load_seed_golden("50", "H33", 0.75)
    assert (int(g["calls"]), int(g["cells"])) == (3288, 28795400)
    g = load_seed_golden("1k", "H33", 0.75)
    assert int(g["calls"]) == 64060


def test_cython_build_matches_python():
    import importlib.util
    from oracle import build_cython
    out = build_cython.build()
    so = [f for f in os.listdir(out) if f.startswith("swalign_cy") and f.endswith(".so")][0]
    spec = importlib.util.spec_from_file_location("swalign_cy", os.path.join(out, so))
    cy = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cy)
    al_c = cy.LocalAlignment(cy.NucleotideScoringMatrix(2, -1))
    al_p = _al()
    rng = random.Random(3)
    for _ in range(50):
        ref = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 200)))
        q = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 30)))
        a, b = al_c.align(ref, q), al_p.align(ref, q)
        assert (a.score, a.r_pos, a.r_end, a.cigar_str, a.identity) == (b.score, b.r_pos, b.r_end, b.cigar_str, b.identity)
