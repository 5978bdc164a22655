"""The skbio / SSW restatement (oracle/ssw_oracle.c): against the one published known answer, degenerate cases, and the host
emulation of the CUDA kernel's lane logic (lamprey_b200/csrc/lsw_ssw.cuh) against it.  PARITY UNPINNED: scikit-bio is absent."""
import json
import os
import random

import numpy as np
import pytest

import emu_lib
from conftest import GOLD, load_reads, schema_path
from lamprey_b200._native import TASK_DT
from oracle import ssw_oracle, seeding_ref

DOC_TARGET = "AAAAAACTCTCTAAACTCACTAAGGCTCTCTACCCCTCTTCAGAGAAGTCGA"
DOC_QUERY = "ACTAAGGCTCTCTACCCCTCTCAGAGA"
_OPS = " MID"


def test_scikit_bio_documentation_example():
    # the example of skbio.alignment.StripedSmithWaterman's own documentation
    a = ssw_oracle.StripedSmithWaterman(DOC_QUERY)(DOC_TARGET)
    assert (a.optimal_alignment_score, a.suboptimal_alignment_score) == (49, 24)
    assert (a.query_begin, a.query_end, a.target_begin, a.target_end_optimal, a.target_end_suboptimal) == (0, 26, 18, 45, 29)
    assert a.cigar == "20M1D7M"
    assert a.aligned_query_sequence == "ACTAAGGCTCTCTACCCCTC-TCAGAGA"
    assert a.aligned_target_sequence == "ACTAAGGCTCTCTACCCCTCTTCAGAGA"
    assert ssw_oracle.align_sequence_skbio(DOC_TARGET, DOC_QUERY, "x") == (18, 46, 1.0, "x")


def test_degenerate_cases():
    d = ssw_oracle.align("AAAA", "CCCC")
    assert d["score"] == 0 and d["ref_begin"] == -1 and d["cigar"] == []
    d = ssw_oracle.align("ACGTACGTAC", "ACGTACGTAC")
    assert (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["cigar"]) == (20, 0, 9, 0, 9, [(10, "M")])
    d = ssw_oracle.align("TTTTACGTNACGTTTT", "ACGTAACGT")      # N scores 0 against everything: bridged without a penalty
    assert d["score"] == 16 and d["cigar"] == [(9, "M")] and d["matches"] == 8
    d = ssw_oracle.align("ACGTACGT" * 3, "ACGTACGT")           # the FIRST column reaching the best score wins
    assert (d["ref_begin"], d["ref_end"]) == (0, 7)


def test_committed_vectors():
    vecs = json.load(open(os.path.join(GOLD, "ssw_vectors.json")))
    assert len(vecs) > 300 and (vecs[0]["target"], vecs[0]["query"]) == (DOC_TARGET, DOC_QUERY)
    for v in vecs:
        d = ssw_oracle.align(v["target"], v["query"])
        got = (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["score2"], d["ref_end2"], d["matches"],
               "".join("%d%s" % x for x in d["cigar"]))
        want = tuple(v[k] for k in ("score", "target_begin", "target_end_optimal", "query_begin", "query_end", "suboptimal_score",
                                    "target_end_suboptimal", "matches", "cigar"))
        assert got == want


def _check_emu_tasks(reads, seqs, tasks, p=(2, -3, 5, 2), packed=True):
    res, cig = emu_lib.ssw_align_tasks(reads, seqs, tasks, *p, packed=packed)
    for t, r in zip(tasks, res):
        d = ssw_oracle.align(reads[t["read"]][t["start"]:t["end"]], seqs[t["query"]], *p)
        if d["score"] == 0:
            assert r["score"] == 0 and r["r_pos"] == -1 and r["n_cigar"] == 0
            continue
        got = (r["score"], r["r_pos"], r["r_end"] - 1, r["q_pos"], r["q_end"] - 1, r["matches"],
               [(int(c >> 2), _OPS[int(c & 3)]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]])
        assert got == (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["matches"], d["cigar"])
        aligned_cols = sum(n for n, _ in d["cigar"])
        assert r["mismatches"] == aligned_cols - d["matches"]


@pytest.mark.parametrize("seed,alphabet,p", [(1, "ACGT", (2, -3, 5, 2)), (2, "AC", (2, -3, 5, 2)), (3, "ACGT", (1, -1, 2, 1)),
                                             (4, "ACG", (3, -2, 6, 1)), (5, "ACGTN", (2, -3, 5, 2)), (6, "A", (2, -3, 5, 2))])
def test_emu_ssw_align_vs_oracle(seed, alphabet, p):
    rng = random.Random(seed)
    q_alpha = alphabet.replace("N", "") or "A"
    seqs = ["".join(rng.choice(q_alpha) for _ in range(n)) for n in (5, 15, 16, 17, 24, 33, 50, 63)]
    reads = []
    for _ in range(24):
        n = rng.randrange(1, 260)
        r = [rng.choice(alphabet) for _ in range(n)]
        for _ in range(3):
            s = seqs[rng.randrange(len(seqs))]
            if n > len(s) + 2:
                pos, k = rng.randrange(0, n - len(s)), 0
                for ch in s:                                    # a noisy copy: deletions, insertions, substitutions
                    x = rng.random()
                    if x < 0.07:
                        continue
                    if x < 0.14 and pos + k < n:
                        r[pos + k] = rng.choice(alphabet); k += 1
                    if pos + k < n:
                        r[pos + k] = ch; k += 1
        reads.append("".join(r))
    tasks = [(i, 0, len(r), q) for i, r in enumerate(reads) for q in range(len(seqs))]
    tasks += [(i, a, b, q) for i, r in enumerate(reads) for q in (1, 4) for a, b in [sorted((rng.randrange(len(r) + 1), rng.randrange(len(r) + 1)))]]
    _check_emu_tasks(reads, seqs, np.array(tasks, dtype=TASK_DT), p)
    _check_emu_tasks(reads, seqs, np.array(tasks, dtype=TASK_DT), p, packed=False)


def test_emu_ssw_wide_band_takes_the_second_tier():
    # one long gap: the box is far from square, banded_sw's first band is wider than the first tier holds
    q = "ACGTTGCAAGGCTTAACCGGATATCGCGTA"
    reads = [q[:15] + "T" * 20 + q[15:], q[:10] + q[22:]]
    p = (2, -1, 3, 1)
    res, _ = emu_lib.ssw_align_tasks(reads, [q], np.array([(0, 0, len(reads[0]), 0), (1, 0, len(reads[1]), 0)], dtype=TASK_DT), *p)
    _check_emu_tasks(reads, [q], np.array([(0, 0, len(reads[0]), 0), (1, 0, len(reads[1]), 0)], dtype=TASK_DT), p)
    hits, cnt = emu_lib.ssw_find_all(reads, [q], 0.5, *p)
    assert cnt["tier2"] >= 1 and res["score"][0] > 0


def _oracle_rows(reads, seqs, thr):
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = ssw_oracle.find_all(r, seqs, thr)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    return np.array(rows, np.int32).reshape(-1, 9), calls, cells


def ssw_matrix(hits):
    cols = ("read", "query", "start", "end", "matches", "score", "q_pos", "q_end", "depth")
    return np.stack([hits[k].astype(np.int32) for k in cols], 1) if len(hits) else np.zeros((0, 9), np.int32)


@pytest.mark.parametrize("thr,packed", [(0.70, True), (0.75, True), (0.70, False)])
def test_emu_ssw_find_all_vs_oracle_and_golden(thr, packed):
    # packed: the forward scan by the s16x2 DPX lanes of lsw_affine.cuh (what the library runs when match * len(primer) <= 127);
    # otherwise the scalar per-task scan
    reads = load_reads("50")[:12]
    seqs = [s for _, s in seeding_ref.PrimerSet(schema_path("H33")).sequences()]
    hits, cnt = emu_lib.ssw_find_all(reads, seqs, thr, packed=packed)
    want, calls, cells = _oracle_rows(reads, seqs, thr)
    assert np.array_equal(ssw_matrix(hits), want) and len(want) > 50
    assert (cnt["tasks"], cnt["cells"]) == (calls, cells)
    gold = np.load(os.path.join(GOLD, "seed_50_H33_%02d_skbio.npz" % round(thr * 100)))
    assert np.array_equal(want, gold["hits"][gold["hits"][:, 0] < 12])


def test_restated_alignSequence_skbio_and_recursion_agree():
    # the Python restatement of lamprey.py:572-645 (skbio branch) over the same C alignment: same hits as the C recursion
    reads = load_reads("50")[:3]
    ps = seeding_ref.PrimerSet(schema_path("H33"))
    for r in reads:
        py = []
        for k, (name, pseq) in enumerate(ps.sequences()):
            def rec(seq, shift, depth):
                s, e, ident, _ = ssw_oracle.align_sequence_skbio(seq, pseq, name)
                if ident >= 0.75 and (e - s) >= len(pseq) * 0.75:
                    py.append((k, s + shift, e + shift))
                    if len(seq[:s]) >= len(pseq):
                        rec(seq[:s], shift, depth + 1)
                    if len(seq[e:]) >= len(pseq):
                        rec(seq[e:], shift + e, depth + 1)
            rec(r, 0, 0)
        h, _ = ssw_oracle.find_all(r, [s for _, s in ps.sequences()], 0.75)
        assert [(int(x["seq"]), int(x["start"]), int(x["end"])) for x in h] == py


def test_emu_stage2_subreads_vs_oracle():
    # stage 2 (lamprey.py:1637-1658): the sub-read, any length, reverse-complemented for 'rev' targets, is the QUERY and the target
    # reference the target
    from lamprey_b200._native import PAIR_DT
    from lamprey_b200 import seeding
    rng = random.Random(9)
    refs = {}
    for fn in ("H3F3A.fa", "HIST1H3B.fa"):
        refs[fn] = open(os.path.join(GOLD, fn)).read().split("\n", 1)[1].replace("\n", "").upper()
    reads = load_reads("50")[:16] + ["ACGT" * 30, "A" * 50, "", "NNNNACGTTGCANNN"]
    r0 = reads[0]
    reads.append(r0[:120] + "N" + r0[121:260] + "RY" + r0[262:400])        # non-ACGT bases INSIDE an alignment: they score 0, not a mismatch
    pairs = []
    for i, r in enumerate(reads):
        pairs.append((i, 0, len(r), 0))
        for _ in range(2):
            a = rng.randrange(0, max(1, len(r) - 40))
            pairs.append((i, a, min(len(r), a + rng.randrange(20, 1200)), rng.randrange(2)))
    pairs = np.array(pairs, dtype=PAIR_DT)
    for fn, ref in refs.items():
        res, cig = emu_lib.ssw_subreads(reads, pairs, ref)
        n_aligned = 0
        for p, r in zip(pairs, res):
            q = reads[p["read"]][p["start"]:p["end"]]
            if p["revcomp"]:
                q = seeding.rc(q)
            d = ssw_oracle.align(ref, q)
            if d["score"] == 0:
                assert r["score"] == 0 and r["r_pos"] == -1
                continue
            n_aligned += 1
            got = (r["score"], r["r_pos"], r["r_end"] - 1, r["q_pos"], r["q_end"] - 1, r["matches"],
                   [(int(c >> 2), _OPS[int(c & 3)]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]])
            assert got == (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["matches"], d["cigar"])
        assert n_aligned > 40
