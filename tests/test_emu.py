"""Kernel lane logic, emulated on the host (liblsw_emu.so), against the oracle.  This is the same
__host__ __device__ code the CUDA kernels run (ScoreLane<NR>, trace_task<NR>); the GPU tier
(test_gpu_parity.py) repeats these checks through the real kernels and the C ABI."""
import json
import os
import random

import numpy as np
import pytest

import emu_lib
from conftest import GOLD, load_reads, schema_path, load_seed_golden, hits_matrix
from lamprey_b200._native import TASK_DT
from oracle import seeding_ref, c_oracle

_OPS = " MID"


def _seqs(schema):
    return [s for _, s in seeding_ref.PrimerSet(schema_path(schema)).sequences()]


@pytest.mark.parametrize("reads_key,schema,thr", [("50", "H33", 0.70), ("50", "H33", 0.75), ("50", "H31", 0.70),
                                                  ("50", "H31", 0.75), ("1k", "H33", 0.75), ("1k", "H31", 0.70)])
def test_emu_find_all_vs_golden(reads_key, schema, thr):
    gold = load_seed_golden(reads_key, schema, thr)
    hits, cnt = emu_lib.find_all(load_reads(reads_key), _seqs(schema), thr)
    assert np.array_equal(hits_matrix(hits), gold["hits"])
    assert (cnt["tasks"], cnt["cells"]) == (int(gold["calls"]), int(gold["cells"]))


def _check_tasks(reads, seqs, tasks, scoring=(2, -1, -1)):
    res, cig = emu_lib.align_tasks(reads, seqs, tasks, *scoring)
    for t, r in zip(tasks, res):
        want = c_oracle.align(reads[t["read"]][t["start"]:t["end"]], seqs[t["query"]], scoring[0], scoring[1],
                              scoring[2], scoring[2])
        got_cigar = [(int(c) >> 2, _OPS[int(c) & 3]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]]
        got = dict(score=int(r["score"]), q_pos=int(r["q_pos"]), q_end=int(r["q_end"]), r_pos=int(r["r_pos"]),
                   r_end=int(r["r_end"]), matches=int(r["matches"]), mismatches=int(r["mismatches"]), cigar=got_cigar)
        assert got == want, (t, reads[t["read"]][t["start"]:t["end"]], seqs[t["query"]])


def test_emu_align_vs_committed_vectors():
    vecs = [v for v in json.load(open(os.path.join(GOLD, "align_vectors.json")))
            if v["query"] and set(v["query"].upper()) <= set("ACGT")]
    assert len(vecs) > 250
    reads = [v["ref"] for v in vecs]
    seqs = sorted(set(v["query"] for v in vecs))
    tasks = np.zeros(len(vecs), TASK_DT)
    for i, v in enumerate(vecs):
        tasks[i] = (i, 0, len(v["ref"]), seqs.index(v["query"]))
    res, cig = emu_lib.align_tasks(reads, seqs, tasks)
    for v, r in zip(vecs, res):
        cigar = "".join("%d%s" % (int(c) >> 2, _OPS[int(c) & 3]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]])
        got = (r["score"], r["q_pos"], r["q_end"], r["r_pos"], r["r_end"], r["matches"], r["mismatches"], cigar)
        assert got == tuple(v[k] for k in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches", "cigar"))


@pytest.mark.parametrize("seed", range(6))
def test_emu_align_random_tie_heavy(seed):
    rng = random.Random(seed)
    alpha = ["AC", "A", "ACG", "ACGT", "AT", "ACGT"][seed]
    reads, seqs, tasks = [], [], []
    for q in range(12):
        seqs.append("".join(rng.choice(alpha) for _ in range(rng.choice([1, 2, 5, 15, 16, 17, 24, 25, 33, 48, 49, 63]))))
    for i in range(40):
        n = rng.choice([0, 1, 3, 15, 16, 17, 63, 64, 65, 127, 128, 129, 300, 700])
        r = [rng.choice(alpha + ("N" if rng.random() < 0.2 else "")) for _ in range(n)]
        if n > 80 and rng.random() < 0.7:         # plant noisy copies
            for _ in range(3):
                s = seqs[rng.randrange(len(seqs))]
                p = rng.randrange(0, n - len(s))
                for k, ch in enumerate(s):
                    if rng.random() > 0.1:
                        r[p + k] = ch
        reads.append("".join(r).lower() if rng.random() < 0.1 else "".join(r))
    for i, r in enumerate(reads):
        for _ in range(6):
            a = rng.randrange(0, len(r) + 1)
            b = rng.randrange(a, len(r) + 1)
            tasks.append((i, a, b, rng.randrange(len(seqs))))
        tasks.append((i, 0, len(r), rng.randrange(len(seqs))))
    _check_tasks(reads, seqs, np.array(tasks, dtype=TASK_DT))


@pytest.mark.parametrize("scoring", [(1, -1, -1), (3, -2, -2), (2, -3, -1), (1, -2, -3), (2, -1, -2)])
def test_emu_align_general_scoring(scoring):
    rng = random.Random(sum(scoring) + 11)
    seqs = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (8, 15, 20, 24, 31, 40)]
    reads = []
    for i in range(30):
        n = rng.randrange(1, 400)
        r = [rng.choice("ACGT") for _ in range(n)]
        s = seqs[rng.randrange(len(seqs))]
        if n > len(s) + 2:
            p = rng.randrange(0, n - len(s))
            for k, ch in enumerate(s):
                if rng.random() > 0.15:
                    r[p + k] = ch
        reads.append("".join(r))
    tasks = np.array([(i, 0, len(r), q) for i, r in enumerate(reads) for q in range(len(seqs))], dtype=TASK_DT)
    _check_tasks(reads, seqs, tasks, scoring)


def test_emu_find_all_edge_reads():
    seqs = _seqs("H33")
    reads = ["", "A", "ACGT" * 3, "N" * 100, seqs[0], seqs[0] * 7, (seqs[4] + "ACGTTGCA") * 30, "acgt" * 50 + seqs[9].lower()]
    hits, cnt = emu_lib.find_all(reads, seqs, 0.7)
    rows = []
    calls = cells = 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, 0.7)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    assert np.array_equal(hits_matrix(hits), np.array(rows, np.int32).reshape(-1, 10))
    assert (cnt["tasks"], cnt["cells"]) == (calls, cells)


def test_emu_rejects_unsupported_scoring():
    with pytest.raises(RuntimeError):
        emu_lib.align_tasks(["ACGT"], ["ACG"], np.array([(0, 0, 4, 0)], dtype=TASK_DT), 2, 0, -1)
    with pytest.raises(RuntimeError):        # 256 * (mismatch + |gap|) would wrap the s16 half of a profile entry
        emu_lib.align_tasks(["ACGT"], ["ACG"], np.array([(0, 0, 4, 0)], dtype=TASK_DT), 2, -200, -1)


def test_emu_most_negative_mismatch_that_fits():
    # mismatch + |gap| = -128 is the edge of the packed profile entry: results must still equal the oracle's
    rng = random.Random(77)
    seqs = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (9, 16, 21)]
    reads = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (40, 150, 333)]
    tasks = np.array([(i, 0, len(r), q) for i, r in enumerate(reads) for q in range(len(seqs))], dtype=TASK_DT)
    _check_tasks(reads, seqs, tasks, (2, -129, -1))


def test_emu_fast_traceback_equals_full_window():
    # the short-window certified traceback must agree with the a-priori exact full-window kernel, and the
    # tie-heavy inputs must exercise the fallback path
    rng = random.Random(5)
    seqs = ["A" * 15, "AC" * 10, "".join(rng.choice("AC") for _ in range(30)), "ACGTTGCAACGTAGCTAGCTAACG"]
    reads = ["A" * 300, "AC" * 200, "".join(rng.choice("AC") for _ in range(900)), "".join(rng.choice("ACGT") for _ in range(500))]
    tasks = np.array([(i, 0, len(r), q) for i, r in enumerate(reads) for q in range(len(seqs))], dtype=TASK_DT)
    st = {}
    fast, cig1 = emu_lib.align_tasks(reads, seqs, tasks, stats=st)
    full, cig2 = emu_lib.align_tasks(reads, seqs, tasks, force_full=True)
    assert st["fallbacks"] >= 0
    for k in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches", "n_cigar"):
        assert np.array_equal(fast[k], full[k]), k
    for a, b in zip(fast, full):
        assert np.array_equal(cig1[a["cigar_off"]:a["cigar_off"] + a["n_cigar"]], cig2[b["cigar_off"]:b["cigar_off"] + b["n_cigar"]])
    _check_tasks(reads, seqs, tasks)


def _oracle_rows(reads, seqs, thr, scoring=(2, -1, -1)):
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr, scoring[0], scoring[1], scoring[2], scoring[2])
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    return np.array(rows, np.int32).reshape(-1, 10), calls, cells


@pytest.mark.parametrize("seed", range(5))
def test_emu_find_all_random_schemas_and_block_edges(seed):
    # random schemas (lengths up to the 63-row limit), reads made of noisy copies so the recursion is deep, read
    # lengths on and around the 64-column block grid the DP-reuse path is built on; compared with the C oracle
    rng = random.Random(100 + seed)
    scoring = [(2, -1, -1), (2, -1, -1), (1, -1, -1), (3, -2, -2), (2, -3, -1)][seed]
    thr = [0.7, 0.75, 0.6, 0.7, 0.8][seed]
    qlens = [rng.choice([5, 15, 16, 17, 20, 21, 24, 25, 28, 31, 33, 40, 47, 56, 57, 63]) for _ in range(7)]
    if scoring[0] == 3:
        qlens = [min(q, 40) for q in qlens]          # match * len + |gap| <= 127
    seqs = ["".join(rng.choice("ACGT") for _ in range(q)) for q in qlens]
    reads = []
    for n in [0, 63, 64, 65, 127, 128, 129, 191, 192, 193, 256, 320, 448, 449, 700, 1024, 1500, 2111, 3000]:
        r = []
        while len(r) < n:
            s = seqs[rng.randrange(len(seqs))]
            if rng.random() < 0.3:
                s = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 40)))
            for ch in s:
                u = rng.random()
                if u < 0.04:
                    continue
                r.append(rng.choice("ACGT") if u < 0.08 else ch)
                if u > 0.96:
                    r.append(rng.choice("ACGT"))
        reads.append("".join(r[:n]))
    want, calls, cells = _oracle_rows(reads, seqs, thr, scoring)
    for reuse in (True, False):
        hits, cnt = emu_lib.find_all(reads, seqs, thr, *scoring, reuse=reuse)
        assert np.array_equal(hits_matrix(hits), want), (seed, reuse)
        assert (cnt["tasks"], cnt["cells"]) == (calls, cells)
    assert len(want) > 50


@pytest.mark.parametrize("scoring", [(7, -5, -5), (5, -12, -2)])
def test_emu_scoring_outside_the_short_window_range(scoring):
    # neighbouring cells may differ by 16 or more: the short-window kernel (H modulo 16) hands every candidate to the
    # full-window kernel
    rng = random.Random(sum(scoring) + 5)
    seqs = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (8, 12, 15, 17)]
    reads = []
    for i in range(40):
        n = rng.randrange(1, 300)
        r = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(3):
            s = seqs[rng.randrange(len(seqs))]
            if n > len(s) + 2:
                p = rng.randrange(0, n - len(s))
                for k, ch in enumerate(s):
                    if rng.random() > 0.1:
                        r[p + k] = ch
        reads.append("".join(r))
    tasks = np.array([(i, 0, len(r), q) for i, r in enumerate(reads) for q in range(len(seqs))], dtype=TASK_DT)
    _check_tasks(reads, seqs, tasks, scoring)
    want, calls, cells = _oracle_rows(reads, seqs, 0.7, scoring)
    hits, cnt = emu_lib.find_all(reads, seqs, 0.7, *scoring)
    assert np.array_equal(hits_matrix(hits), want)
    assert (cnt["tasks"], cnt["cells"]) == (calls, cells) and cnt["fallbacks"] == cnt["traced"] > 0


def test_emu_long_sparse_reads_use_second_level_block_bests():
    # few primer copies in long reads: child slices span thousands of columns, so the merge takes whole groups of block
    # bests from the second level (lsw_reuse.cuh); slice ends fall on and around the 1024-column group grid
    rng = random.Random(77)
    seqs = ["".join(rng.choice("ACGT") for _ in range(q)) for q in (18, 22, 24, 31, 50)]
    reads = []
    for n in (1023, 1024, 1025, 2047, 2048, 2049, 3500, 5000, 7000, 9000, 4096, 6144):
        r = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(rng.randrange(2, 7)):
            sq = seqs[rng.randrange(len(seqs))]
            p = rng.choice([0, 1024 - len(sq), 1024, 2048 - 5, n - len(sq), rng.randrange(0, n - len(sq))])
            p = max(0, min(n - len(sq), p))
            for k, ch in enumerate(sq):
                if rng.random() > 0.06:
                    r[p + k] = ch
        reads.append("".join(r))
    want, calls, cells = _oracle_rows(reads, seqs, 0.7)
    for reuse in (True, False):
        hits, cnt = emu_lib.find_all(reads, seqs, 0.7, 2, -1, -1, reuse=reuse)
        assert np.array_equal(hits_matrix(hits), want), reuse
        assert (cnt["tasks"], cnt["cells"]) == (calls, cells)
    assert len(want) > 20


def test_emu_recursion_deeper_than_pythons_limit():
    # exact tandem repeats of a short sequence: the greedy recursion finds one copy per level (swalign keeps the LAST maximum, so
    # everything else is in the left child; SSW keeps the FIRST, so it is in the right child) -- 1500 levels, where the
    # reference's Python would raise RecursionError at ~1000.  Both paths, against the C oracles.
    from oracle import ssw_oracle
    read, seqs = "ACGTA" * 1500, ["ACGTA", "TACGT"]
    want, st = c_oracle.find_all(read, seqs, 0.7)
    want = want[np.argsort(want["start"], kind="stable")]
    hits, cnt = emu_lib.find_all([read], seqs, 0.7)
    assert len(hits) == len(want) >= 2000 and cnt["tasks"] == st["calls"] and st["max_depth"] == 1499
    for a, b in (("query", "seq"), ("start", "start"), ("end", "end"), ("depth", "depth"), ("matches", "matches")):
        assert (hits[a] == want[b]).all()
    want, st = ssw_oracle.find_all(read, seqs, 0.7)
    want = want[np.argsort(want["start"], kind="stable")]
    hits, cnt = emu_lib.ssw_find_all([read], seqs, 0.7)
    assert len(hits) == len(want) and cnt["tasks"] == st["calls"] and st["max_depth"] >= 1499
    for a, b in (("query", "seq"), ("start", "start"), ("end", "end"), ("depth", "depth"), ("matches", "matches")):
        assert (hits[a] == want[b]).all()


def test_emu_stops_at_the_librarys_round_limit():
    # a 1-base sequence on a homopolymer recurses once per base; past 30 000 levels the library (and its emulation) report an
    # internal error instead of wrapping the 16-bit depth field (limits)
    with pytest.raises(RuntimeError, match="rc=-6"):
        emu_lib.find_all(["A" * 30500], ["A"], 0.7)
