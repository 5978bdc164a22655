"""Random small batches of the DEFAULT (skbio / SSW) seeding path through the C ABI against oracle/ssw_oracle.c
(run on a B200: python tools/fuzz_gpu_ssw.py [iterations [first seed]]).  Several scorings (those with match * len > 127 take the
scalar forward scan, the others the packed one), reads of 0..3000 bases with lower case and non-ACGT bytes, planted noisy copies
with insertions and deletions so that the banded traceback and both band tiers are exercised; plus random stage-2 sub-reads."""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
from lamprey_b200 import _native          # noqa: E402
from oracle import ssw_oracle             # noqa: E402

COLS = ("read", "query", "start", "end", "matches", "score", "q_pos", "q_end", "depth")
_COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
eng = _native.Engine(0)
bad = 0
limited = 0
for it in range(iters):
    rng = random.Random(seed0 + it)
    p = rng.choice([(2, -3, 5, 2), (2, -3, 5, 2), (1, -1, 2, 1), (3, -2, 6, 1), (2, -1, 3, 1), (2, -4, 4, 4), (1, -3, 5, 2)])
    seqs = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 64))) for _ in range(rng.randrange(1, 7))]
    thr = rng.choice([0.5, 0.6, 0.7, 0.75, 0.8, 0.9])
    reads = []
    for _ in range(rng.randrange(1, 30)):
        n = rng.choice([0, 1, 5, 15, 16, 17, 63, 64, 65, 200, 700, 1500, 3000])
        r = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(rng.randrange(0, 8)):
            s = seqs[rng.randrange(len(seqs))]
            if n > len(s) + 4:
                pos, k = rng.randrange(0, n - len(s) - 4), 0
                for ch in s:
                    x = rng.random()
                    if x < 0.05:
                        continue
                    if x < 0.10 and pos + k < n:
                        r[pos + k] = rng.choice("ACGT"); k += 1
                    if pos + k < n:
                        r[pos + k] = ch; k += 1
        for _ in range(rng.randrange(0, 4)):
            if n:
                r[rng.randrange(n)] = rng.choice("NnacgtRY-")
        reads.append("".join(r))
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = ssw_oracle.find_all(r.upper(), seqs, thr, *p)       # the engine upper-cases reads (lamprey feeds upper-case FASTQ)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    want = np.array(rows, np.int32).reshape(-1, 9)
    eng.set_ssw(*p)
    eng.set_queries(seqs)
    eng.upload_read_list(reads)
    hits = eng.find_all(thr)
    got = np.stack([hits[k].astype(np.int32) for k in COLS], 1) if len(hits) else np.zeros((0, 9), np.int32)
    c = eng.last_counters
    ok = np.array_equal(got, want) and (c["tasks"], c["cells"]) == (calls, cells)
    # stage 2: random sub-reads of the same batch against one of the reads as "reference"
    refs = [r for r in reads if 50 <= len(r) <= 700]
    if ok and refs:
        ref = refs[0].upper()
        pairs = []
        for _ in range(20):
            i = rng.randrange(len(reads))
            if not reads[i]:
                continue
            a = rng.randrange(len(reads[i])); b = min(len(reads[i]), a + rng.randrange(1, 900))
            pairs.append((i, a, b, rng.randrange(2)))
        if pairs:
            # scorings with cheap gaps put random sequences into the linear regime of local alignment (alignments with hundreds of
            # gap bases): the stage-2 kernel then reports its band limit instead of guessing.  Counted, not a mismatch.
            try:
                res, cig = eng.ssw_subreads(np.array(pairs, dtype=_native.PAIR_DT), ref)
            except ValueError as e:
                if "limits" not in str(e):
                    raise
                limited += 1
                res, cig, pairs = [], [], []
            for (i, a, b, rcf), r in zip(pairs, res):
                q = reads[i][a:b].upper()
                if rcf:
                    q = "".join(_COMP.get(ch, ch) for ch in reversed(q))
                d = ssw_oracle.align(ref, q, *p)
                if d["score"] == 0:
                    ok = ok and r["score"] == 0
                else:
                    runs = [(int(x >> 2), " MID"[int(x & 3)]) for x in cig[r["cigar_off"]:r["cigar_off"] + r["n_cigar"]]]
                    ok = ok and (r["score"], r["r_pos"], r["r_end"] - 1, r["q_pos"], r["q_end"] - 1, runs) == \
                        (d["score"], d["ref_begin"], d["ref_end"], d["read_begin"], d["read_end"], d["cigar"])
    if not ok:
        bad += 1
        print("MISMATCH iteration", it, p, thr, [len(s) for s in seqs], got.shape, want.shape, (c["tasks"], calls), (c["cells"], cells))
print("fuzz (skbio path): %d iterations, %d mismatches, %d stage-2 batches beyond the band limit" % (iters, bad, limited))
sys.exit(1 if bad else 0)
