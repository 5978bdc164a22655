"""Random small batches through the C ABI against the C oracle (run on a B200: python tools/fuzz_gpu.py [iterations [first seed]]).
Odd shapes on purpose: 1..6 sequences of 1..63 bases, reads of 0..5000 bases with lower case and non-ACGT bytes,
thresholds 0.5..0.95, several scorings, planted noisy copies so the recursion is exercised."""
import random
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import hits_matrix          # noqa: E402
from lamprey_b200 import _native          # noqa: E402
from oracle import c_oracle               # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = _native.Engine(0)
bad = 0
for it in range(iters):
    rng = random.Random(seed0 + it)
    M, X, gap = rng.choice([(2, -1, -1), (2, -1, -1), (1, -1, -1), (3, -2, -2), (2, -3, -1), (1, -2, -3), (2, -1, -2), (4, -3, -2)])
    g = -gap
    qmax = min(63, (127 - g) // M)
    seqs = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(1, qmax + 1))) for _ in range(rng.randrange(1, 7))]
    thr = rng.choice([0.5, 0.6, 0.7, 0.75, 0.8, 0.9, 0.95])
    reads = []
    for _ in range(rng.randrange(1, 40)):
        n = rng.choice([0, 1, 5, 31, 32, 33, 64, 200, 700, 1500, 5000])
        r = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(rng.randrange(0, 8)):
            s = seqs[rng.randrange(len(seqs))]
            if n > len(s):
                p = rng.randrange(0, n - len(s))
                for k, ch in enumerate(s):
                    if rng.random() > 0.1:
                        r[p + k] = ch
        for _ in range(rng.randrange(0, 4)):
            if n:
                r[rng.randrange(n)] = rng.choice("NnacgtRY-")
        reads.append("".join(r))
    rows, calls, cells = [], 0, 0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr, M, X, gap, gap)
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    want = np.array(rows, np.int32).reshape(-1, 10)
    eng.set_scoring(M, X, gap, gap)
    eng.set_queries(seqs)
    eng.upload_read_list(reads)
    got = hits_matrix(eng.find_all(thr))
    c = eng.last_counters
    ok = np.array_equal(got, want) and (c["tasks"], c["cells"]) == (calls, cells)
    if not ok:
        bad += 1
        print("MISMATCH iteration", it, (M, X, g), thr, [len(s) for s in seqs], got.shape, want.shape, (c["tasks"], calls), (c["cells"], cells))
print("fuzz: %d iterations, %d mismatches" % (iters, bad))
sys.exit(1 if bad else 0)
