"""GPU tests of the library-level pool (include/lsw.h "Pool"): one batch split over several contexts / GPUs with the
host-side gather, double-buffered streaming, and the 16-byte hit records.  Everything is compared byte for byte with the
single-context result, which test_gpu_parity.py pins to the oracle and the goldens."""
import numpy as np
import pytest

from conftest import load_reads, schema_path, load_seed_golden, hits_matrix
from lamprey_b200 import _native, seeding, synth

pytestmark = pytest.mark.gpu


def _seqs(schema):
    return [s for _, s in seeding.PrimerSet(schema_path(schema)).sequences()]


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_compact_hits_equal_wide_hits(engine):
    engine.set_scoring(2, -1, -1, -1)
    engine.set_queries(_seqs("H33"))
    engine.upload_read_list(load_reads("1k"))
    engine.set_overlap(4)
    try:
        wide = engine.find_all(0.70)
        compact = engine.fetch_hits_compact(len(wide))
    finally:
        engine.set_overlap(None)
    assert compact.dtype.itemsize == 16 and len(compact) == len(wide) > 10000
    back = _native.expand_hits(compact, 2, -1, -1)
    assert back.tobytes() == wide.tobytes()              # every field, score and survivor flag included
    c = engine.last_counters
    assert c["device_bytes"] > 0 and c["reuse_dropped"] == 0 and 0 < c["ms_score_round0"] <= c["ms_score"]
    assert 0 < c["computed_cells_round0"] <= c["computed_cells"]


@pytest.mark.parametrize("compact,pieces", [(False, 1), (True, 1), (True, 3)])
def test_pool_split_over_contexts_equals_one_context(compact, pieces):
    # ONE batch dealt over several contexts (every visible GPU; on a single-GPU box two contexts of GPU 0 stand in for two
    # devices) with the host-side gather == the same batch on one context, byte for byte
    devices = list(range(_n_gpus())) if _n_gpus() > 1 else [0, 0]
    concat, offs = synth.generate(600, seed=21, mean_len=1500)
    seqs = _seqs("H33")
    one = _native.Engine(0)
    one.set_scoring(2, -1, -1, -1); one.set_queries(seqs); one.upload_reads(concat, offs)
    one.set_overlap(4)
    want = one.find_all(0.7)
    want_cnt = one.last_counters
    one.close()
    pool = _native.Pool(devices, 2 if pieces > 1 else 1)
    pool.configure(seqs, 0.7, allowed_overlap=4, compact=compact, pieces=pieces)      # pieces: ranges per device, alternating contexts
    out = np.zeros(len(want) + 10, pool.hit_dtype)
    hits, cnt, times = pool.wait(pool.submit(concat, offs, out))
    if compact:
        hits = _native.expand_hits(hits)
    assert hits.tobytes() == want.tobytes()
    assert (cnt["tasks"], cnt["cells"], cnt["hits"]) == (want_cnt["tasks"], want_cnt["cells"], want_cnt["hits"])
    assert sum(times["reads"]) == 600 and sum(times["hits"]) == len(want) and min(times["reads"][:len(devices)]) > 0
    # a buffer that is too small: reported with the needed count, nothing written past it
    small = np.zeros(100, pool.hit_dtype)
    with pytest.raises(OverflowError) as e:
        pool.wait(pool.submit(concat, offs, small))
    assert e.value.needed == len(want)
    pool.close()


def test_seed_reads_devices_equals_golden():
    devices = list(range(_n_gpus())) if _n_gpus() > 1 else [0, 0]
    reads = load_reads("1k")
    ps = seeding.PrimerSet(schema_path("H33"))
    gold = load_seed_golden("1k", "H33", 0.75)
    batch = seeding.seed_reads(reads, ps, 0.75, devices=devices)
    assert np.array_equal(hits_matrix(batch.hits), gold["hits"])
    assert np.array_equal(batch.coverage, gold["coverage"])
    assert (batch.counters["tasks"], batch.counters["cells"]) == (int(gold["calls"]), int(gold["cells"]))


def test_seed_stream_double_buffered_equals_per_batch_seeding(engine):
    ps = seeding.PrimerSet(schema_path("H33"))
    batches = [synth.generate(n, seed=40 + k, mean_len=900) for k, n in enumerate((64, 1, 200, 0, 33))]
    got = list(seeding.seed_stream(batches, ps, 0.7, devices=[0], depth=2))
    assert len(got) == len(batches)
    for (concat, offs), b in zip(batches, got):
        want = seeding.seed_reads((concat, offs), ps, 0.7)
        assert b.hits.tobytes() == want.hits.tobytes() and len(b) == len(offs) - 1


def test_pool_errors():
    pool = _native.Pool([0])
    with pytest.raises(ValueError):                       # not configured yet
        pool.submit(np.zeros(4, np.uint8), np.array([0, 4]), np.zeros(8, _native.HIT_DT))
    with pytest.raises(ValueError):
        pool.configure(["ACGN"], 0.7)
    with pytest.raises(NotImplementedError):
        pool.configure(["ACGT"], 0.7, gap=-2, gap_ext=-1)
    pool.configure(["ACGTACGTAC"], 0.7)
    hits, cnt, _ = pool.wait(pool.submit(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(8, _native.HIT_DT)))
    assert len(hits) == 0 and cnt["tasks"] == 0          # an empty batch
    pool.close()
    with pytest.raises(RuntimeError):
        _native.Pool([99])


def test_device_subreads_and_processLamplicon_seed_stage():
    # stage 1 of processLamplicon behind the batched call: device sub-reads == the host mirror of extractAmpliconAroundTarget
    from lamprey_b200 import chain, pipeline
    reads = load_reads("1k")
    ps = seeding.PrimerSet(schema_path("H33"))
    for kw in (dict(), dict(args=__import__("types").SimpleNamespace(swalign=False))):
        seeds = pipeline.seed_batch(reads, ps, 0.70, **kw)
        n_sub = 0
        for i in range(len(reads)):
            al = seeds.alignments(i, pruned=True)
            want = chain.target_subreads(ps, al, len(reads[i]), i) if al else []
            assert seeds.target_subreads(i) == want
            n_sub += len(want)
        assert n_sub == len(seeds.subreads) and n_sub > (500 if not kw else 50)
        k = next(i for i in range(len(reads)) if len(seeds.target_subreads(i)) >= 2)
        st = pipeline.processLamplicon_seed_stage(seeds, k, 7, reads[k], ps, 0.70)
        assert st.target_depth == sum(1 for a in st.alignments if a.primer_name in ("T", "Tc")) == len(st.subread_slices)
        sid, direction, sub = st.subread_slices[1]
        s, e, d = st.alignment_interval_dict[sid]
        assert sid == "7_1" and sub == reads[k][s:e] and d == direction
    eng = _native.Engine(0)
    eng.set_scoring(2, -1, -1, -1); eng.set_queries(_seqs("H33")); eng.upload_read_list(reads[:20])
    eng.find_all(0.7)
    with pytest.raises(ValueError):                       # the hits carry no survivor flags
        eng.chain_subreads([0, 1], [1, 0], 0, 1)
    eng.close()


def test_skbio_path_over_several_contexts_equals_golden():
    # the default (StripedSmithWaterman) seeding branch through the pool: one batch split over contexts, gathered in read order
    devices = list(range(_n_gpus())) if _n_gpus() > 1 else [0, 0]
    reads = load_reads("1k")
    ps = seeding.PrimerSet(schema_path("H33"))
    gold = np.load(__import__("os").path.join(__import__("conftest").GOLD, "seed_1k_H33_70_skbio.npz"))
    batch = seeding.seed_reads(reads, ps, 0.70, devices=devices, method="skbio")
    cols = ("read", "query", "start", "end", "matches", "score", "q_pos", "q_end", "depth")
    got = np.stack([batch.hits[k].astype(np.int32) for k in cols], 1)
    assert np.array_equal(got, gold["hits"])
    assert (batch.counters["tasks"], batch.counters["cells"]) == (int(gold["calls"]), int(gold["cells"]))
    one = seeding.seed_reads(reads, ps, 0.70, method="skbio")
    assert batch.hits.tobytes() == one.hits.tobytes() and np.array_equal(batch.identity, one.identity)


def test_fastq_files_to_hits_through_the_library(tmp_path):
    # FASTQ files -> hits with nothing but library code in between: the multi-threaded scanner (lsw_fastq_*), page-locked text and
    # hit buffers, the pool's double buffering, the per-read index (lsw_hits_index); against the golden seeding result
    reads = load_reads("1k")
    ps = seeding.PrimerSet(schema_path("H33"))
    gold = load_seed_golden("1k", "H33", 0.75)
    cuts = [0, 1, 400, 400, 1000]                          # one file holds a single read, one is empty
    paths = []
    for k in range(len(cuts) - 1):
        p = tmp_path / ("f%d.fastq" % k)
        with open(p, "w") as fp:
            for i in range(cuts[k], cuts[k + 1]):
                fp.write("@read%d ch=%d\n%s\n+\n%s\n" % (i, i, reads[i], "I" * len(reads[i])))
        paths.append(str(p))
    G = hits_matrix_gold = gold["hits"]
    for pinned in (False, True):
        got = list(seeding.seed_fastq_files(paths, ps, 0.75, devices=[0], pinned=pinned))
        assert [p for p, _ in got] == paths
        for k, (_, b) in enumerate(got):
            want = G[(G[:, 0] >= cuts[k]) & (G[:, 0] < cuts[k + 1])].copy()
            want[:, 0] -= cuts[k]
            assert len(b) == cuts[k + 1] - cuts[k] and np.array_equal(hits_matrix(b.hits), want)
            assert np.array_equal(b.coverage, gold["coverage"][cuts[k]:cuts[k + 1]])
    # one file as one batch, parsed into page-locked memory, hits downloaded into a caller-owned page-locked buffer
    whole = tmp_path / "all.fastq"
    with open(whole, "w") as fp:
        for i, r in enumerate(reads):
            fp.write("@read%d\n%s\n+\n%s\n" % (i, r, "#" * len(r)))
    buf = _native.PinnedArray(len(G) + 100, _native.HIT16_DT)
    batch, names = seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], pinned=True, hit_buffer=buf.array, with_names=True)
    assert np.array_equal(hits_matrix(batch.hits), G) and np.array_equal(batch.coverage, gold["coverage"])
    assert names[999] == "read999" and len(names) == 1000
    assert batch[3][1] == float(gold["coverage"][3]) and len(batch[3][0]) == int((G[:, 0] == 3).sum())
    # a caller's buffer that is too small is replaced for that batch, not an error
    tiny = _native.PinnedArray(10, _native.HIT16_DT)
    again = seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], hit_buffer=tiny.array)
    assert again.hits.tobytes() == batch.hits.tobytes()
    with pytest.raises(ValueError):
        seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], hit_buffer=np.zeros(10, np.int32))
    buf.close(); tiny.close()
