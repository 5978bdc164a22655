"""Parity tests proper: the CUDA kernels, called through the C ABI (liblsw.so) and the drop-in
Python surface, against the oracle and the committed golden fixtures.  Bit-exact: all outputs are
integers (scores, coordinates, CIGAR runs, match counts); identity is derived in Python doubles."""
import io
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLD, load_reads, schema_path, load_seed_golden, hits_matrix
from lamprey_b200 import seeding, swalign, synth
from lamprey_b200._native import TASK_DT
from oracle import seeding_ref, c_oracle, swalign_ref

pytestmark = pytest.mark.gpu
_OPS = " MID"


def _seqs(schema):
    return [s for _, s in seeding_ref.PrimerSet(schema_path(schema)).sequences()]


def test_readme_vector_through_dropin_surface():
    sw = swalign.LocalAlignment(swalign.NucleotideScoringMatrix(2, -1))       # lamprey.py:1860-1865
    a = sw.align('ACACACTA', 'AGCACACA')
    assert (a.score, a.cigar_str, a.matches, a.mismatches) == (12, '1M1I5M1D1M', 7, 2)
    assert (a.q_pos, a.q_end, a.r_pos, a.r_end) == (0, 8, 0, 8)
    assert a.identity == 7 / 9
    buf = io.StringIO()
    a.dump(out=buf)
    ref = io.StringIO()
    swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1)).align('ACACACTA', 'AGCACACA').dump(out=ref)
    assert buf.getvalue() == ref.getvalue()
    z = sw.align('AAAA', 'CCCC')
    assert (z.score, z.q_pos, z.q_end, z.r_pos, z.r_end, z.cigar, z.identity) == (0, 4, 4, 4, 4, [], 0)
    e = sw.align('', 'ACGT')
    assert (e.score, e.q_pos, e.r_pos, e.r_end) == (0, 0, 0, 0)


@pytest.mark.parametrize("reads_key,schema,thr", [("50", "H33", 0.70), ("50", "H33", 0.75), ("50", "H31", 0.70),
                                                  ("50", "H31", 0.75), ("1k", "H33", 0.70), ("1k", "H33", 0.75),
                                                  ("1k", "H31", 0.70), ("1k", "H31", 0.75)])
def test_find_all_vs_golden(engine, reads_key, schema, thr):
    gold = load_seed_golden(reads_key, schema, thr)
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(_seqs(schema))
    engine.upload_read_list(load_reads(reads_key))
    hits = engine.find_all(thr)
    assert np.array_equal(hits_matrix(hits), gold["hits"])
    c = engine.last_counters
    assert (c["tasks"], c["cells"]) == (int(gold["calls"]), int(gold["cells"]))


def test_seed_reads_dropin_matches_reference_function(engine):
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    batch = seeding.seed_reads(reads, ps, 0.75)
    gold = load_seed_golden("50", "H33", 0.75)
    assert np.array_equal(batch.coverage, gold["coverage"])
    al = swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1))
    for i in (3, 11):
        want, wcov = seeding_ref.find_all_primer_alignments(al, reads[i], seeding_ref.PrimerSet(schema_path("H33")), 0.75)
        got, gcov = batch[i]
        assert [(h.start, h.end, h.identity, h.primer_name) for h in got] == [h.astuple() for h in want]
        assert gcov == wcov
    # function-level drop-ins (same signatures as lamprey.py:596, 699)
    sw = swalign.LocalAlignment(swalign.NucleotideScoringMatrix(2, -1))
    got, cov = seeding.findAllPrimerAlignments(sw, reads[3], ps, 0.75, None)
    want, wcov = seeding_ref.find_all_primer_alignments(al, reads[3], seeding_ref.PrimerSet(schema_path("H33")), 0.75)
    assert [(h.start, h.end, h.identity, h.primer_name) for h in got] == [h.astuple() for h in want] and cov == wcov
    t = ps.primer_dict["T"]
    got = seeding.findPrimerAlignments(sw, reads[11], t, "T", 0.7, None)
    want = seeding_ref.find_primer_alignments(al, reads[11], t, "T", 0.7)
    assert [(h.start, h.end, h.identity) for h in got] == [(h.start, h.end, h.identity) for h in want]


def _check_tasks(engine, reads, seqs, tasks, scoring=(2, -1, -1)):
    engine.set_scoring(scoring[0], scoring[1], scoring[2], scoring[2])
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    res, cig = engine.align_tasks(tasks)
    for t, r in zip(tasks, res):
        want = c_oracle.align(reads[t["read"]][t["start"]:t["end"]], seqs[t["query"]], scoring[0], scoring[1],
                              scoring[2], scoring[2])
        got_cigar = [(int(c) >> 2, _OPS[int(c) & 3]) for c in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]]
        got = dict(score=int(r["score"]), q_pos=int(r["q_pos"]), q_end=int(r["q_end"]), r_pos=int(r["r_pos"]),
                   r_end=int(r["r_end"]), matches=int(r["matches"]), mismatches=int(r["mismatches"]), cigar=got_cigar)
        assert got == want, (t, reads[t["read"]][t["start"]:t["end"]], seqs[t["query"]])


def test_align_many_vs_committed_vectors(engine):
    vecs = [v for v in json.load(open(os.path.join(GOLD, "align_vectors.json")))
            if v["query"] and set(v["query"].upper()) <= set("ACGT")]
    reads = [v["ref"] for v in vecs]
    seqs = sorted(set(v["query"] for v in vecs))
    iv = np.array([(i, 0, len(v["ref"]), seqs.index(v["query"])) for i, v in enumerate(vecs)])
    batch = swalign.align_many(reads, seqs, iv)
    for i, v in enumerate(vecs):
        a = batch[i]
        got = (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end, a.matches, a.mismatches, a.cigar_str, a.identity)
        assert got == tuple(v[k] for k in ("score", "q_pos", "q_end", "r_pos", "r_end", "matches", "mismatches",
                                           "cigar", "identity"))


@pytest.mark.parametrize("seed", range(6))
def test_align_random_tie_heavy(engine, seed):
    rng = random.Random(seed)
    alpha = ["AC", "A", "ACG", "ACGT", "AT", "ACGT"][seed]
    seqs = ["".join(rng.choice(alpha) for _ in range(rng.choice([1, 2, 5, 15, 16, 17, 24, 25, 33, 48, 49, 63])))
            for _ in range(12)]
    reads, tasks = [], []
    for i in range(60):
        n = rng.choice([0, 1, 3, 15, 16, 17, 63, 64, 65, 127, 128, 129, 300, 700, 2500])
        r = [rng.choice(alpha + ("N" if rng.random() < 0.2 else "")) for _ in range(n)]
        if n > 80 and rng.random() < 0.7:
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
    _check_tasks(engine, reads, seqs, np.array(tasks, dtype=TASK_DT))


@pytest.mark.parametrize("scoring", [(1, -1, -1), (3, -2, -2), (2, -3, -1), (1, -2, -3), (2, -1, -2)])
def test_align_general_scoring(engine, scoring):
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
    _check_tasks(engine, reads, seqs, tasks, scoring)


def test_find_all_edge_reads(engine):
    seqs = _seqs("H33")
    reads = ["", "A", "ACGT" * 3, "N" * 100, seqs[0], seqs[0] * 7, (seqs[4] + "ACGTTGCA") * 30, "acgt" * 50 + seqs[9].lower()]
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    hits = engine.find_all(0.7)
    rows = []
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, 0.7)
        h = h[np.argsort(h["start"], kind="stable")]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    assert np.array_equal(hits_matrix(hits), np.array(rows, np.int32).reshape(-1, 10))
    with pytest.raises(ZeroDivisionError):
        seeding.seed_reads(reads, seeding.PrimerSet(schema_path("H33")), 0.7)[0]     # empty read: lamprey.py:731


@pytest.mark.parametrize("cfg,n,schema", [("synth_2kb_8pct", 96, "H33"), ("synth_5kb_12pct", 40, "H31"),
                                          ("synth_20_50kb_12pct", 6, "H33")])
def test_find_all_synthetic_vs_oracle(engine, cfg, n, schema):
    # the BASELINE.json synthetic configs at sizes the C oracle finishes in seconds
    kw = dict(synth.CONFIGS[cfg])
    kw["schema"] = schema
    concat, offs = synth.generate(n, **kw)
    seqs = _seqs(schema)
    if cfg == "synth_20_50kb_12pct":
        seqs = seqs[:40]
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.upload_reads(concat, offs)
    hits = engine.find_all(0.7)
    reads = synth.reads_as_strings(concat, offs)
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, 0.7)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    assert np.array_equal(hits_matrix(hits), np.array(rows, np.int32).reshape(-1, 10))
    assert (engine.last_counters["tasks"], engine.last_counters["cells"]) == (calls, cells)


def test_full_size_properties(engine):
    # size-independent properties at a size the oracle cannot check directly (65536 reads, ~1.3e8 bases):
    # (1) sharding invariance: seeding two halves separately gives the same hits as the whole batch;
    # (2) hits of one (read, sequence) never overlap and lie inside the read;
    # (3) every hit passes the reference predicate in Python doubles;
    # (4) idempotence: a second pass returns identical bytes.
    concat, offs = synth.generate(65536, workers=8, **synth.CONFIGS["synth_2kb_8pct"])
    seqs = _seqs("H33")
    qlen = np.array([len(s) for s in seqs])
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.upload_reads(concat, offs)
    whole = engine.find_all(0.7)
    again = engine.find_all(0.7)
    assert whole.tobytes() == again.tobytes()
    half = len(offs) // 2
    engine.upload_reads(concat[:offs[half]], offs[:half + 1])
    a = engine.find_all(0.7).copy()
    engine.upload_reads(concat[offs[half]:], offs[half:] - offs[half])
    b = engine.find_all(0.7).copy()
    b["read"] += half
    assert np.concatenate([a, b]).tobytes() == whole.tobytes()
    lens = np.diff(offs)
    assert (whole["start"] >= 0).all() and (whole["end"] <= lens[whole["read"]]).all() and (whole["start"] < whole["end"]).all()
    order = np.lexsort((whole["start"], whole["query"], whole["read"]))
    s = whole[order]
    same = (s["read"][1:] == s["read"][:-1]) & (s["query"][1:] == s["query"][:-1])
    assert (s["start"][1:][same] >= s["end"][:-1][same]).all()
    m, x = whole["matches"].astype(np.int64), whole["mismatches"].astype(np.int64)
    assert (m / (m + x) >= 0.7).all() and ((whole["end"] - whole["start"]) >= qlen[whole["query"]] * 0.7).all()
    assert (whole["score"] == 2 * m - x).all()


def test_two_contexts_split_one_batch(engine):
    # a batch split over two contexts (own CUDA streams, own host threads) with lsw_set_read_base must give
    # exactly the hits of the unsplit batch -- the way bench.py's e2e leg overlaps copies with kernels
    import threading
    from lamprey_b200 import _native
    concat, offs = synth.generate(3000, seed=21, mean_len=900)
    seqs = _seqs("H33")
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.set_read_base(0)
    engine.upload_reads(concat, offs)
    whole = engine.find_all(0.7).copy()
    other = _native.Engine(0)
    other.set_scoring(2, -1, -1, -1)
    other.set_queries(seqs)
    shards = seeding.shard_reads(offs, 2)
    parts = [None, None]

    def work(k, eng):
        lo, hi = shards[k]
        eng.set_read_base(lo)
        eng.upload_reads(concat[offs[lo]:offs[hi]], offs[lo:hi + 1] - offs[lo])
        parts[k] = eng.find_all(0.7).copy()

    ths = [threading.Thread(target=work, args=(k, e)) for k, e in enumerate((engine, other))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    engine.set_read_base(0)
    other.close()
    assert np.concatenate(parts).tobytes() == whole.tobytes()


def test_device_overlap_pruning(engine):
    from lamprey_b200 import chain
    reads = load_reads("1k")
    ps = seeding.PrimerSet(schema_path("H33"))
    batch = seeding.seed_reads(reads, ps, 0.70, allowed_overlap=4)
    gold = load_seed_golden("1k", "H33", 0.70)
    assert np.array_equal(hits_matrix(batch.hits), gold["hits"])           # pruning only sets the flag
    assert 0 < batch.kept.sum() < len(batch.hits)
    for i in range(0, len(reads), 7):
        al = batch.alignments(i)
        if al:
            want = chain.removeOverlappingPrimerAlignments(False, al, 4)  # lamprey.py:1536
            got = batch.alignments(i, pruned=True)
            assert [(a.start, a.end, a.primer_name) for a in got] == [(a.start, a.end, a.primer_name) for a in want]


@pytest.mark.parametrize("seed", range(5))
def test_find_all_random_schemas_and_block_edges(engine, seed, monkeypatch):
    # random schemas / scoring / thresholds, deep recursion, read lengths on and around the 64-column block grid of
    # the DP-reuse path; with and without the reuse, against the C oracle
    rng = random.Random(100 + seed)
    scoring = [(2, -1, -1), (2, -1, -1), (1, -1, -1), (3, -2, -2), (2, -3, -1)][seed]
    thr = [0.7, 0.75, 0.6, 0.7, 0.8][seed]
    qlens = [rng.choice([5, 15, 16, 17, 20, 21, 24, 25, 28, 31, 33, 40, 47, 56, 57, 63]) for _ in range(7)]
    if scoring[0] == 3:
        qlens = [min(q, 40) for q in qlens]
    seqs = ["".join(rng.choice("ACGT") for _ in range(q)) for q in qlens]
    reads = []
    for n in [0, 63, 64, 65, 127, 128, 129, 191, 192, 193, 256, 320, 448, 449, 700, 1024, 1500, 2111, 3000] * 3:
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
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr, scoring[0], scoring[1], scoring[2], scoring[2])
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    want = np.array(rows, np.int32).reshape(-1, 10)
    engine.set_scoring(scoring[0], scoring[1], scoring[2], scoring[2])
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    for no_reuse in ("", "1"):
        if no_reuse:
            monkeypatch.setenv("LSW_NO_REUSE", "1")
        else:
            monkeypatch.delenv("LSW_NO_REUSE", raising=False)
        hits = engine.find_all(thr)
        assert np.array_equal(hits_matrix(hits), want), (seed, no_reuse)
        c = engine.last_counters
        assert (c["tasks"], c["cells"]) == (calls, cells)
        if no_reuse:
            assert c["computed_cells"] == c["cells"]
        else:
            assert c["computed_cells"] <= c["cells"]
    engine.set_scoring(2, -1, -1, -1)


def test_error_behaviour_of_the_engine(engine):
    # scoring and queries may arrive in either order; the pair is checked when work is submitted
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(["ACGT" * 15])
    engine.set_scoring(3, -2, -2, -2)                   # 3 * 60 + 2 > 127: accepted here ...
    engine.upload_read_list(["ACGT" * 40])
    with pytest.raises(ValueError, match="127"):        # ... refused when it would have to be computed
        engine.find_all(0.7)
    engine.set_queries(["ACGT" * 5])                    # now it fits
    assert len(engine.find_all(0.7)) > 0
    with pytest.raises(ValueError):
        engine.set_queries(["ACGN"])                    # non-ACGT query
    with pytest.raises(ValueError):
        engine.set_queries(["A" * 64])                  # longer than the register file holds
    with pytest.raises(ValueError):
        engine.set_scoring(2, 0, -1, -1)                # mismatch must be negative
    with pytest.raises(NotImplementedError):
        engine.set_scoring(2, -1, -2, -1)               # affine gaps
    with pytest.raises(ValueError, match="-128"):
        engine.set_scoring(2, -200, -1, -1)             # 256 * (mismatch + |gap|) would wrap its s16 half
    engine.set_scoring(2, -127, -1, -1)                 # the most negative mismatch that fits
    with pytest.raises(ValueError):
        engine.find_all(0.0)                            # the reference would recurse forever
    engine.set_scoring(2, -1, -1, -1)


@pytest.mark.parametrize("scoring", [(7, -5, -5), (5, -12, -2)])
def test_scoring_outside_the_short_window_range(engine, scoring):
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
    _check_tasks(engine, reads, seqs, tasks, scoring)
    rows = []
    for i, r in enumerate(reads):
        h, _ = c_oracle.find_all(r, seqs, 0.7, scoring[0], scoring[1], scoring[2], scoring[2])
        h = h[np.argsort(h["start"], kind="stable")]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    engine.set_scoring(scoring[0], scoring[1], scoring[2], scoring[2])
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    hits = engine.find_all(0.7)
    assert np.array_equal(hits_matrix(hits), np.array(rows, np.int32).reshape(-1, 10))
    assert engine.last_counters["fallbacks"] == engine.last_counters["traced"] > 0
    engine.set_scoring(2, -1, -1, -1)


def test_long_sparse_reads_use_second_level_block_bests(engine, monkeypatch):
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
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, 0.7, 2, -1, -1, -1)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    want = np.array(rows, np.int32).reshape(-1, 10)
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    for l1 in ("1", "0"):                       # read by lsw_upload_reads
        monkeypatch.setenv("LSW_BB_L1", l1)
        engine.upload_read_list(reads)
        hits = engine.find_all(0.7)
        assert np.array_equal(hits_matrix(hits), want), l1
        assert (engine.last_counters["tasks"], engine.last_counters["cells"]) == (calls, cells)
    assert len(want) > 20


def _oracle_chunk(args):
    reads, seqs, thr, base = args
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]
        cells += st["cells"]
        rows += [(base + i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"],
                  x["depth"]) for x in h]
    return rows, calls, cells


def test_find_all_1024_synthetic_reads_vs_oracle(engine):
    # hits, call count and cell count cross-checked against the oracle on >= 1000 reads of the
    # headline workload (the C oracle runs one process per host core)
    import multiprocessing as mp
    n, thr = 1024, 0.7
    concat, offs = synth.generate(n, **synth.CONFIGS["synth_2kb_8pct"])
    seqs = _seqs("H33")
    reads = synth.reads_as_strings(concat, offs)
    cores = max(1, min(32, os.cpu_count() or 1))
    step = (n + cores - 1) // cores
    with mp.get_context("fork").Pool(cores) as pool:
        parts = pool.map(_oracle_chunk, [(reads[k:k + step], seqs, thr, k) for k in range(0, n, step)])
    rows = [r for p in parts for r in p[0]]
    calls, cells = sum(p[1] for p in parts), sum(p[2] for p in parts)
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.upload_reads(concat, offs)
    hits = engine.find_all(thr)
    assert np.array_equal(hits_matrix(hits), np.array(rows, np.int32).reshape(-1, 10))
    assert (engine.last_counters["tasks"], engine.last_counters["cells"]) == (calls, cells)
    assert len(rows) > 100000


@pytest.mark.parametrize("wl", ["synth_5kb_12pct", "synth_20_50kb_12pct"])
def test_find_all_1000_reads_of_the_large_configs_vs_committed_digest(engine, wl):
    # BASELINE.json configs[3] and [4] (5 kb / 12 %; 20-50 kb with the 40-sequence schema): 1000-read slices, oracle results
    # committed as counts + SHA-256 of the int32 hit matrix (tests/golden/make_golden_more.py)
    import hashlib
    gold = json.load(open(os.path.join(GOLD, "synth_digests.json")))[wl]
    concat, offs = synth.generate(gold["reads"], **synth.CONFIGS[wl])
    assert int(offs[-1]) == gold["bases"]
    ps = seeding.PrimerSet(schema_path("H33"))
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries([s for _, s in synth.schema_sequences(ps, wl)])
    engine.upload_reads(concat, offs)
    hits = engine.find_all(gold["threshold"])
    mat = hits_matrix(hits)
    assert mat[:8].tolist() == gold["first_rows"]
    assert (len(mat), engine.last_counters["tasks"], engine.last_counters["cells"]) == (gold["hits"], gold["calls"], gold["cells"])
    assert hashlib.sha256(np.ascontiguousarray(mat, dtype="<i4").tobytes()).hexdigest() == gold["sha256"]
    assert engine.last_counters["fallbacks"] == 0


def test_seed_fastq_one_file_one_batch(engine, tmp_path):
    # FASTQ file -> seeds in one call (the online mode's one file = one batch), against the golden seeding result
    reads = load_reads("50")
    path = tmp_path / "batch.fastq"
    with open(path, "w") as fp:
        for i, r in enumerate(reads):
            fp.write("@read%d runid=x\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
    batch, names = seeding.seed_fastq(str(path), seeding.PrimerSet(schema_path("H33")), 0.75, with_names=True)
    gold = load_seed_golden("50", "H33", 0.75)
    assert np.array_equal(hits_matrix(batch.hits), gold["hits"])
    assert names[:2] == ["read0", "read1"] and len(batch) == len(reads)


def test_hit_buffer_overflow_is_reported_and_retried(engine):
    # a hit buffer that is too small: the C ABI reports LSW_E_OVERFLOW with the number of hits needed (nothing is
    # written past the buffer), the Python shim enlarges the buffer and runs again
    import ctypes
    from lamprey_b200 import _native
    reads = load_reads("50")
    seqs = _seqs("H33")
    gold = load_seed_golden("50", "H33", 0.75)
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(seqs)
    engine.upload_read_list(reads)
    lib, h = engine._lib, engine._h
    assert lib.lsw_set_hit_capacity(h, 16) == 0
    n = ctypes.c_int64(0)
    cnt = np.zeros(1, _native.COUNTERS_DT)
    rc = lib.lsw_find_all(h, 0.75, ctypes.byref(n), cnt.ctypes.data_as(ctypes.c_void_p))
    assert rc == _native.LSW_E_OVERFLOW and n.value == len(gold["hits"])
    hits = engine.find_all(0.75)                       # still 16: overflows once more, then retried with room
    assert np.array_equal(hits_matrix(hits), gold["hits"])
    assert lib.lsw_set_hit_capacity(h, 0) == 0         # back to the default sizing
