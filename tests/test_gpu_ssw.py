"""GPU parity of LAMPrey's DEFAULT seeding path (scikit-bio StripedSmithWaterman semantics, lamprey.py:572-593 + 596-645 with
args.swalign False): the CUDA kernels of lamprey_b200/csrc/lsw_ssw.cuh, through the C ABI and the drop-in Python surface,
against the restated oracle (oracle/ssw_oracle.c) and the committed goldens.  Bit-exact (all integers)."""
import json
import os
import random
import sys
import types

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_reads, schema_path
from lamprey_b200 import seeding, skbio_ssw, synth
from lamprey_b200._native import TASK_DT
from oracle import ssw_oracle, seeding_ref

pytestmark = pytest.mark.gpu
_OPS = " MID"
COLS = ("read", "query", "start", "end", "matches", "score", "q_pos", "q_end", "depth")


def _seqs(schema):
    return [s for _, s in seeding_ref.PrimerSet(schema_path(schema)).sequences()]


def ssw_matrix(hits):
    return np.stack([hits[k].astype(np.int32) for k in COLS], 1) if len(hits) else np.zeros((0, 9), np.int32)


@pytest.mark.parametrize("reads_key,schema,thr", [("50", "H33", 0.70), ("50", "H33", 0.75), ("50", "H31", 0.75),
                                                  ("1k", "H33", 0.70), ("1k", "H33", 0.75), ("1k", "H31", 0.70), ("1k", "H31", 0.75)])
def test_ssw_find_all_vs_golden(engine, reads_key, schema, thr):
    gold = np.load(os.path.join(GOLD, "seed_%s_%s_%02d_skbio.npz" % (reads_key, schema, round(thr * 100))))
    engine.set_ssw(2, -3, 5, 2)
    engine.set_queries(_seqs(schema))
    engine.upload_read_list(load_reads(reads_key))
    hits = engine.find_all(thr)
    assert np.array_equal(ssw_matrix(hits), gold["hits"])
    c = engine.last_counters
    assert (c["tasks"], c["cells"]) == (int(gold["calls"]), int(gold["cells"]))
    engine.set_scoring(2, -1, -1, -1)


def test_skbio_documentation_example_through_the_dropin_class():
    a = skbio_ssw.StripedSmithWaterman("ACTAAGGCTCTCTACCCCTCTCAGAGA")("AAAAAACTCTCTAAACTCACTAAGGCTCTCTACCCCTCTTCAGAGAAGTCGA")
    assert (a.optimal_alignment_score, a.query_begin, a.query_end, a.target_begin, a.target_end_optimal, a.cigar) == \
           (49, 0, 26, 18, 45, "20M1D7M")
    assert a.aligned_query_sequence == "ACTAAGGCTCTCTACCCCTC-TCAGAGA" and a.aligned_target_sequence == "ACTAAGGCTCTCTACCCCTCTTCAGAGA"
    with pytest.raises(NotImplementedError):
        skbio_ssw.StripedSmithWaterman("ACGT", gap_open_penalty=3)
    long_q = "ACGTTGCA" * 12                                            # longer than the register rows: the stage-2 kernel takes it
    b = skbio_ssw.StripedSmithWaterman(long_q)("TT" + long_q[:50] + "G" + long_q[50:] + "AA")
    w = ssw_oracle.StripedSmithWaterman(long_q)("TT" + long_q[:50] + "G" + long_q[50:] + "AA")
    assert (b.optimal_alignment_score, b.target_begin, b.target_end_optimal, b.cigar) == \
           (w.optimal_alignment_score, w.target_begin, w.target_end_optimal, w.cigar)
    # the module boundary of lamprey.py:14-17: ./lib first on sys.path, then `from skbio.alignment import StripedSmithWaterman`
    sys.path.insert(0, os.path.join(ROOT, "lib"))
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "skbio" or k.startswith("skbio.")}
        from skbio.alignment import StripedSmithWaterman
        assert StripedSmithWaterman is skbio_ssw.StripedSmithWaterman
    finally:
        for k in [k for k in sys.modules if k == "skbio" or k.startswith("skbio.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path.remove(os.path.join(ROOT, "lib"))


def test_ssw_align_many_vs_committed_vectors():
    vecs = [v for v in json.load(open(os.path.join(GOLD, "ssw_vectors.json"))) if 0 < len(v["query"]) <= 63 and v["target"]]
    queries = sorted(set(v["query"] for v in vecs))
    targets = sorted(set(v["target"] for v in vecs))
    qi = {q: i for i, q in enumerate(queries)}
    ti = {t: i for i, t in enumerate(targets)}
    res = skbio_ssw.align_many(queries, targets, intervals=[(ti[v["target"]], 0, len(v["target"]), qi[v["query"]]) for v in vecs])
    assert len(res) > 300
    for v, a in zip(vecs, res):
        if v["score"] == 0:
            assert a.optimal_alignment_score == 0 and a.target_begin == -1
            continue
        assert (a.optimal_alignment_score, a.target_begin, a.target_end_optimal, a.query_begin, a.query_end, a.matches, a.cigar) == \
               (v["score"], v["target_begin"], v["target_end_optimal"], v["query_begin"], v["query_end"], v["matches"], v["cigar"])


def test_dropin_functions_take_the_skbio_branch_like_the_reference():
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    args = types.SimpleNamespace(swalign=False)           # the reference's default (lamprey.py:2297)
    got, cov = seeding.findAllPrimerAlignments(None, reads[7], ps, 0.75, args)
    seqs = ps.sequences()
    h, _ = ssw_oracle.find_all(reads[7], [s for _, s in seqs], 0.75)
    h = h[np.argsort(h["start"], kind="stable")]
    want = [(int(x["start"]), int(x["end"]), float(x["matches"]) / float(len(seqs[x["seq"]][1])), seqs[x["seq"]][0]) for x in h]
    assert [(a.start, a.end, a.identity, a.primer_name) for a in got] == want and len(want) > 3
    assert cov == float(sum(e - s for s, e, _, _ in want)) / float(len(reads[7]))
    t = ps.primer_dict["T"]
    pre = seeding.findPrimerAlignments(None, reads[7], t, "T", 0.7, args)
    h, _ = ssw_oracle.find_all(reads[7], [t], 0.7)          # the C recursion emits in pre-order
    assert [(a.start, a.end, a.identity) for a in pre] == [(int(x["start"]), int(x["end"]), float(x["matches"]) / len(t)) for x in h]
    one = seeding.alignSequence_skbio(reads[7][:400], t, "T")
    assert (one.start, one.end, one.identity, one.primer_name) == ssw_oracle.align_sequence_skbio(reads[7][:400], t, "T")
    # and with args.swalign True the swalign branch
    sw_args = types.SimpleNamespace(swalign=True)
    from lamprey_b200 import swalign
    al = swalign.LocalAlignment(swalign.NucleotideScoringMatrix(2, -1))
    a1, _ = seeding.findAllPrimerAlignments(al, reads[7], ps, 0.75, sw_args)
    a2, _ = seeding.findAllPrimerAlignments(al, reads[7], ps, 0.75, None)
    assert [(a.start, a.end, a.identity) for a in a1] == [(a.start, a.end, a.identity) for a in a2]


@pytest.mark.parametrize("seed", range(4))
def test_ssw_random_tie_heavy_batches_vs_oracle(engine, seed):
    rng = random.Random(100 + seed)
    alphabet = ["ACGT", "AC", "ACGTN", "ACG"][seed]
    q_alpha = alphabet.replace("N", "")
    seqs = ["".join(rng.choice(q_alpha) for _ in range(rng.randrange(4, 64))) for _ in range(7)]
    reads = []
    for _ in range(60):
        n = rng.randrange(0, 700)
        r = [rng.choice(alphabet) for _ in range(n)]
        for _ in range(5):
            s = seqs[rng.randrange(len(seqs))]
            if n > len(s) + 2:
                pos, k = rng.randrange(0, n - len(s)), 0
                for ch in s:
                    x = rng.random()
                    if x < 0.06:
                        continue
                    if x < 0.12 and pos + k < n:
                        r[pos + k] = rng.choice(alphabet); k += 1
                    if pos + k < n:
                        r[pos + k] = ch; k += 1
        reads.append("".join(r))
    thr = [0.7, 0.6, 0.75, 0.5][seed]
    engine.set_ssw(2, -3, 5, 2)
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    hits = engine.find_all(thr)
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = ssw_oracle.find_all(r, seqs, thr)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    assert np.array_equal(ssw_matrix(hits), np.array(rows, np.int32).reshape(-1, 9)) and len(rows) > 20
    assert (engine.last_counters["tasks"], engine.last_counters["cells"]) == (calls, cells)
    engine.set_scoring(2, -1, -1, -1)


def test_ssw_synthetic_concatemers_vs_oracle(engine):
    concat, offs = synth.generate(48, seed=31, mean_len=1500)
    reads = synth.reads_as_strings(concat, offs)
    seqs = _seqs("H33")
    batch = seeding.seed_reads((concat, offs), seeding.PrimerSet(schema_path("H33")), 0.7, method="skbio")
    rows = []
    for i, r in enumerate(reads):
        h, _ = ssw_oracle.find_all(r, seqs, 0.7)
        h = h[np.argsort(h["start"], kind="stable")]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    assert np.array_equal(ssw_matrix(batch.hits), np.array(rows, np.int32).reshape(-1, 9)) and len(rows) > 2000
    k = len(rows) // 2
    assert batch.identity[k] == rows[k][4] / len(seqs[rows[k][1]])


def test_stage2_subreads_against_the_target_reference(engine):
    # lamprey.py:1637-1669 behind one batched call: seeds -> device sub-reads -> every sub-read (rc for 'rev') against the target
    # reference, compared with the oracle alignment of the same strings
    from lamprey_b200 import pipeline
    reads = load_reads("1k")[:300]
    ps = seeding.PrimerSet(schema_path("H33"))
    ref = open(os.path.join(GOLD, "H3F3A.fa")).read().split("\n", 1)[1].replace("\n", "").upper()
    seeds = pipeline.seed_batch(reads, ps, 0.70)
    recs = pipeline.stage2_align_batch(reads, seeds, ref)
    assert len(recs) == len(seeds.subreads) > 150
    for x, rec in zip(seeds.subreads, recs):
        q = reads[x["read"]][x["start"]:x["end"]]
        if x["direction"] == 1:
            q = seeding.rc(q)
        a = ssw_oracle.StripedSmithWaterman(q)(ref)
        if a.optimal_alignment_score == 0:
            assert rec["score"] == 0 and rec["reference_start"] == -1
            continue
        assert (rec["score"], rec["reference_start"], rec["cigar"]) == (a.optimal_alignment_score, a.target_begin, a.cigar)
        assert rec["query_sequence"] == a.aligned_query_sequence.replace('-', '')
    # the drop-in class takes the same path for a long query
    q = reads[5][:700]
    a = skbio_ssw.StripedSmithWaterman(q)(ref)
    b = ssw_oracle.StripedSmithWaterman(q)(ref)
    assert (a.optimal_alignment_score, a.target_begin, a.target_end_optimal, a.query_begin, a.query_end, a.cigar) == \
           (b.optimal_alignment_score, b.target_begin, b.target_end_optimal, b.query_begin, b.query_end, b.cigar)
    assert a.aligned_query_sequence == b.aligned_query_sequence and a.aligned_target_sequence == b.aligned_target_sequence
