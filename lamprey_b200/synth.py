"""Synthetic ONT-like LAMP concatemer reads.

A LAMP product is a concatemer of alternating forward / reverse-complement copies of the amplicon.
Template of a read: optional ONT adapter (+ barcode) at the head, then copies of
``amplicon + rc(amplicon)`` entered at a random phase, cut to the drawn length; then sequencing
errors (substitution / insertion / deletion in equal parts) at the requested rate.
Everything is vectorised numpy (``numpy.random.Generator(PCG64(seed))``), generated in chunks of
reads so that 10^6 reads of 2 kb take seconds, not minutes.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_GOLD = os.path.join(os.path.dirname(_HERE), "tests", "golden")

_ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b

# the 141 bp F2..B2c stretch of the H3F3A target reference (data/references/H3F3A.fa[41:182]) and
# the HIST1H3B target; read from the committed fixtures so the GPU box does not need /root/reference
def _fasta(name):
    with open(os.path.join(_GOLD, name)) as fp:
        return "".join(l.strip() for l in fp if not l.startswith(">")).upper()


def amplicon(schema="H33"):
    if schema == "H33":
        return _fasta("H3F3A.fa")[41:182]
    return _fasta("HIST1H3B.fa")[200:380]


ONT_RAP_TOP = "GGCGTCTGCTTGGGTGTTTAACCTTTTTTTTTTAATGTACTTCGTTCAGTTACGTATTGCT"


def _rc_bytes(a):
    return _COMP[a[::-1]]


def draw_lengths(rng, n, kind="gamma", mean=2000, lo=200, hi=20000):
    if kind == "gamma":
        ln = rng.gamma(4.0, mean / 4.0, n)
    else:
        ln = rng.uniform(lo, hi, n)
    return np.clip(ln, lo, hi).astype(np.int64)


def _chunk(args):
    (n, seed, ci, mean_len, error, schema, length_kind, min_len, max_len, adapter_p, heads) = args
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, ci])))
    amp = np.frombuffer(amplicon(schema).encode(), np.uint8)
    unit = np.concatenate([amp, _rc_bytes(amp)])
    U = len(unit)
    hl = max(len(h) for h in heads)
    head_tab = np.zeros((len(heads), hl), np.uint8)
    head_len = np.array([len(h) for h in heads], np.int64)
    for k, h in enumerate(heads):
        head_tab[k, :len(h)] = np.frombuffer(h, np.uint8)
    L = draw_lengths(rng, n, length_kind, mean_len, min_len, max_len)
    phase = rng.integers(0, U, n)
    has_head = rng.random(n) < adapter_p
    which = rng.integers(0, len(heads), n)
    hlen = np.where(has_head, head_len[which], 0)
    starts = np.zeros(n + 1, np.int64)
    np.cumsum(L, out=starts[1:])
    tot = int(starts[-1])
    rid = np.repeat(np.arange(n, dtype=np.int32), L)
    pos = np.arange(tot, dtype=np.int64) - starts[rid]
    body = pos - hlen[rid]
    tpl = unit[(phase[rid] + np.maximum(body, 0)) % U]
    in_head = body < 0
    if in_head.any():
        tpl[in_head] = head_tab[which[rid[in_head]], pos[in_head]]
    r = rng.random(tot, dtype=np.float32)
    is_sub = r < error / 3
    is_ins = (r >= error / 3) & (r < 2 * error / 3)
    is_del = (r >= 2 * error / 3) & (r < error)
    rnd = _ACGT[rng.integers(0, 4, tot, dtype=np.uint8)]
    tpl = np.where(is_sub, rnd, tpl)
    counts = np.ones(tot, np.uint8)
    counts[is_ins] = 2
    counts[is_del] = 0
    out = np.repeat(tpl, counts)
    ends = np.cumsum(counts, dtype=np.int64)
    ins_at = ends[is_ins] - 1                       # the second copy of an inserted position is random
    out[ins_at] = _ACGT[rng.integers(0, 4, len(ins_at), dtype=np.uint8)]
    new_len = np.add.reduceat(counts.astype(np.int64), starts[:-1]) if tot else np.zeros(n, np.int64)
    return out, new_len


def _mem_available():
    try:
        with open("/proc/meminfo") as fp:
            for line in fp:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


JOB_BYTES_PER_BASE = 56        # peak of _chunk: int32 read ids, three int64 index arrays, float32 draws, masks, the two texts


def workers_that_fit(workers, chunk, mean_read_len, mem_share):
    """The worker count capped so that the jobs in flight stay within `mem_share` of the host's available memory."""
    avail = _mem_available()
    if avail is None or not mem_share:
        return workers
    per_job = JOB_BYTES_PER_BASE * chunk * mean_read_len
    return int(max(1, min(workers, (avail * mem_share) // max(1, per_job))))


def generate(n_reads, seed=3, mean_len=2000, error=0.08, schema="H33", length_kind="gamma", min_len=200,
             max_len=20000, adapter_p=0.5, heads=None, chunk=16384, workers=1, mem_share=None):
    """-> (uint8 ASCII concat, int64 offsets[n+1]).  heads: optional list of byte strings, one of which
    (uniformly) replaces the ONT adapter at the read head (config 5 injects barcode + adapter).
    Each chunk of reads has its own PCG64 stream seeded by (seed, chunk index), so the result does
    not depend on `workers`.  workers > 1 forks a pool: call it BEFORE initialising CUDA.  mem_share: fraction of the host's
    available memory the jobs in flight may take (several ranks generate at once on a multi-GPU box); fewer workers run if more
    would not fit."""
    if heads is None:
        heads = [ONT_RAP_TOP.encode()]
    if workers > 1 and mem_share:
        workers = workers_that_fit(workers, min(chunk, n_reads), mean_len if length_kind == "gamma" else (min_len + max_len) / 2.0, mem_share)
    jobs = [(min(chunk, n_reads - c0), seed, ci, mean_len, error, schema, length_kind, min_len, max_len, adapter_p,
             heads) for ci, c0 in enumerate(range(0, n_reads, chunk))]
    if workers > 1 and len(jobs) > 1:
        # a ProcessPoolExecutor, not a multiprocessing.Pool: if the kernel's OOM killer takes a worker, the executor raises
        # BrokenProcessPool, where Pool.map would wait for ever
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(min(workers, len(jobs)), mp_context=mp.get_context("fork")) as pool:
            parts = list(pool.map(_chunk, jobs))
    else:
        parts = [_chunk(j) for j in jobs]
    concat = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint8)
    lens = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int64)
    offs = np.zeros(n_reads + 1, np.int64)
    np.cumsum(lens, out=offs[1:])
    return concat, offs


_BARCODES = ["AAGAAAGTTGTCGGTGTCTTTGTG", "TCGATTCCGTTTGTAGTCGTCTGT", "GAGTCTTGTGTCCCAGTTACCAGG",
             "TTCGGATTCTATCGTGTTTCCCTA", "CTTGTCCAGGGTTTGTGTAACCTT", "TTCTCGCAAAGGCAGAAAGTAGTC",
             "GTGTTACCGTGGGAATGAATCCTT", "TTCAGGGAACAAACCAAGTTACGT", "AACTAGGCACAGCGAGTCTTGGTT"]   # ontBC01..09

CONFIGS = {
    # BASELINE.json configs[2..4] / SURVEY.md section 8d
    "synth_2kb_8pct": dict(mean_len=2000, error=0.08, min_len=200, max_len=20000, seed=3, schema="H33"),
    # (chunk: reads per generator job.  A job holds ~50 bytes per base while it runs: 16384 reads of 35 kb were 27 GB per worker
    # process, and a dozen workers per rank on a multi-GPU box ran the host out of memory)
    "synth_5kb_12pct": dict(mean_len=5000, error=0.12, min_len=500, max_len=50000, seed=4, schema="H33", chunk=4096),
    # long-read stress: dense concatemers, one adapter + random barcode at the read head
    "synth_20_50kb_12pct": dict(length_kind="uniform", min_len=20000, max_len=50000, error=0.12, seed=5,
                                schema="H33", adapter_p=1.0, chunk=2048,
                                heads=[(ONT_RAP_TOP + b).encode() for b in _BARCODES]),
}
# schema used with a workload: number of NAMED sequences taken from the schema file (None = all)
SCHEMA_NAMES = {"synth_20_50kb_12pct": 20}


def schema_sequences(primer_set, workload):
    """The (name, sequence) search list of a workload: BASELINE.json's long-read config uses the first 20 named
    sequences of the H3.3 schema (40 with reverse complements), the others the whole schema."""
    pairs = primer_set.sequences()
    n = SCHEMA_NAMES.get(workload)
    if n is None:
        return pairs
    keep = set(list(primer_set.primer_dict)[:n])
    return [(nm, sq) for nm, sq in pairs if (nm[:-1] if nm.endswith("c") and nm[:-1] in primer_set.primer_dict else nm) in keep]


def reads_as_strings(concat, offs):
    b = concat.tobytes()
    return [b[offs[i]:offs[i + 1]].decode("ascii") for i in range(len(offs) - 1)]
