"""CPU tests of the Python layers above the pool (seeding.seed_reads(devices=...), seed_stream, seed_fastq, seed_fastq_files,
caller-owned hit buffers) with the device replaced by the host emulation of the kernels: `_native.Pool` is swapped for a stand-in
that seeds every submitted batch with emu_lib.find_all and answers through the same submit / wait protocol (16-byte records,
OverflowError with the needed count), `_native.PinnedArray` for plain memory.  The real pool and the real page-locked buffers run in
tests/test_gpu_pool.py (-m gpu) and, for the pool's own threading, tests/test_pool_host.py."""
import numpy as np
import pytest

import emu_lib
from conftest import load_reads, schema_path, load_seed_golden, hits_matrix
from lamprey_b200 import _native, seeding


def pack16(w):
    c = np.zeros(len(w), _native.HIT16_DT)
    c["read"] = w["read"]; c["start"] = w["start"]; c["span"] = w["end"] - w["start"]; c["query"] = w["query"]
    c["depth"] = w["depth"]; c["matches"] = w["matches"] | (w["reserved"] << 7); c["mismatches"] = w["mismatches"]
    c["q_pos"] = w["q_pos"]; c["q_end"] = w["q_end"]
    return c


class EmuPool(object):
    instances = 0

    def __init__(self, devices=(0,), contexts_per_device=1):
        EmuPool.instances += 1
        self.devices, self.hit_format, self.scoring, self._done, self._next = list(devices), _native.HITS_WIDE, (2, -1, -1), {}, 1

    def configure(self, seqs, threshold, match=2, mismatch=-1, gap=-1, gap_ext=-1, allowed_overlap=None, compact=False, pieces=1, ssw=None):
        assert ssw is None
        self.seqs, self.thr, self.scoring = list(seqs), threshold, (match, mismatch, gap)
        self.overlap = -1 if allowed_overlap is None else allowed_overlap
        self.hit_format = _native.HITS_COMPACT if compact else _native.HITS_WIDE

    @property
    def hit_dtype(self):
        return _native.HIT16_DT if self.hit_format == _native.HITS_COMPACT else _native.HIT_DT

    def submit(self, ascii_u8, offs, out):
        assert out.dtype == self.hit_dtype
        b = np.asarray(ascii_u8).tobytes()
        reads = [b[offs[i]:offs[i + 1]].decode() for i in range(len(offs) - 1)]
        hits, cnt = emu_lib.find_all(reads, self.seqs, self.thr, *self.scoring, allowed_overlap=self.overlap) if reads else (np.zeros(0, _native.HIT_DT), dict(tasks=0, cells=0))
        rec = pack16(hits) if self.hit_format == _native.HITS_COMPACT else hits
        if len(rec) <= len(out):
            out[:len(rec)] = rec
        self._done[self._next] = (out, len(rec), cnt)
        self._next += 1
        return self._next - 1

    def wait(self, ticket):
        out, n, cnt = self._done.pop(ticket)
        if n > len(out):
            e = OverflowError("hit buffer too small: %d records are needed" % n)
            e.needed = n
            raise e
        return out[:n], dict(cnt), {"total_s": 0.0}

    def close(self):
        pass


class PlainArray(object):
    def __init__(self, count, dtype):
        self.array = np.zeros(int(count), dtype)

    def close(self):
        self.array = None


@pytest.fixture
def emu_device(monkeypatch):
    monkeypatch.setattr(_native, "Pool", EmuPool)
    monkeypatch.setattr(_native, "PinnedArray", PlainArray)
    monkeypatch.setattr(seeding, "_pools", {})
    return EmuPool


def test_fastq_files_to_hits_through_the_python_layers(emu_device, tmp_path):
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    gold = load_seed_golden("50", "H33", 0.75)
    G = gold["hits"]
    cuts = [0, 1, 20, 20, 50]                              # one file holds a single read, one is empty
    paths = []
    for k in range(len(cuts) - 1):
        p = tmp_path / ("f%d.fastq" % k)
        with open(p, "w") as fp:
            for i in range(cuts[k], cuts[k + 1]):
                fp.write("@read%d ch=%d\n%s\n+\n%s\n" % (i, i, reads[i], "I" * len(reads[i])))
        paths.append(str(p))
    for pinned in (False, True):
        got = list(seeding.seed_fastq_files(paths, ps, 0.75, devices=[0], pinned=pinned))
        assert [p for p, _ in got] == paths
        for k, (_, b) in enumerate(got):
            want = G[(G[:, 0] >= cuts[k]) & (G[:, 0] < cuts[k + 1])].copy()
            want[:, 0] -= cuts[k]
            assert len(b) == cuts[k + 1] - cuts[k] and np.array_equal(hits_matrix(b.hits), want)
            assert np.array_equal(b.coverage, gold["coverage"][cuts[k]:cuts[k + 1]])
    whole = tmp_path / "all.fastq"
    with open(whole, "w") as fp:
        for i, r in enumerate(reads):
            fp.write("@read%d\n%s\n+\n%s\n" % (i, r, "#" * len(r)))
    buf = _native.PinnedArray(len(G) + 100, _native.HIT16_DT)
    batch, names = seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], pinned=True, hit_buffer=buf.array, with_names=True)
    assert np.array_equal(hits_matrix(batch.hits), G) and np.array_equal(batch.coverage, gold["coverage"])
    assert names[49] == "read49" and len(names) == 50
    assert batch[3][1] == float(gold["coverage"][3]) and len(batch[3][0]) == int((G[:, 0] == 3).sum())
    tiny = _native.PinnedArray(10, _native.HIT16_DT)       # too small: replaced for that batch, not an error
    again = seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], hit_buffer=tiny.array)
    assert again.hits.tobytes() == batch.hits.tobytes()
    with pytest.raises(ValueError):
        seeding.seed_fastq(str(whole), ps, 0.75, devices=[0], hit_buffer=np.zeros(10, np.int32))
    assert EmuPool.instances <= 2                          # one pool per (devices, depth), reused


def test_seed_reads_devices_overlap_flag_and_stream_overflow(emu_device):
    reads = load_reads("50")
    ps = seeding.PrimerSet(schema_path("H33"))
    gold = load_seed_golden("50", "H33", 0.70)
    for compact in (True, False):
        b = seeding.seed_reads(reads, ps, 0.70, devices=[0, 1], allowed_overlap=4, compact=compact)
        assert np.array_equal(hits_matrix(b.hits), gold["hits"])
        for i in (0, 7, 49):                               # the survivor flag travels through the 16-byte records
            al = b.alignments(i, pruned=True)
            from lamprey_b200 import chain
            want = chain.removeOverlappingPrimerAlignments(False, b.alignments(i), 4)
            assert [(a.start, a.end, a.primer_name) for a in al] == [(a.start, a.end, a.primer_name) for a in want]
        if compact:                                        # the per-read counts of processLamplicon for the whole batch at once
            td, no = b.target_depth(), b.num_ont_seqs(0.70)
            for i in range(len(reads)):
                al = b.alignments(i, pruned=True)
                assert td[i] == sum(1 for a in al if a.primer_name in ('T', 'Tc'))
                assert no[i] == sum(1 for a in al if a.primer_name.startswith('ont') and a.identity >= 0.70)
            assert td.sum() > 50 and (b.target_depth(pruned=False) >= td).all()
    # seed_stream: a batch whose hits do not fit the default buffer (bases / 8 + 4096 records) is seeded again with room
    dense = ["ACGTA" * 1000] * 14                          # a five-base sequence on its own tandem repeat: a hit every five bases
    class P(object):
        primer_dict = {"T": "ACGTA"}
        def sequences(self):
            return [("T", "ACGTA")]
    out = list(seeding.seed_stream([dense, reads[:3]], P(), 0.7, devices=[0]))
    assert len(out[0].hits) == 14 * 1000 > (14 * 5000) // 8 + 4096
    assert (np.diff(out[0].offsets) == 1000).all() and np.allclose(out[0].coverage, 1.0)
    assert len(out[1]) == 3
