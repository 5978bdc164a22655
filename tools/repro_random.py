import sys, random, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import emu_lib
from oracle import c_oracle
seeds = [int(x) for x in sys.argv[1].split(",")]; mult = int(sys.argv[2])
eng = None
if len(sys.argv) > 3:
    from lamprey_b200 import _native
    eng = _native.Engine(0)
for seed in seeds:
    rng = random.Random(100 + seed)
    scoring = [(2, -1, -1), (2, -1, -1), (1, -1, -1), (3, -2, -2), (2, -3, -1)][seed]
    thr = [0.7, 0.75, 0.6, 0.7, 0.8][seed]
    qlens = [rng.choice([5, 15, 16, 17, 20, 21, 24, 25, 28, 31, 33, 40, 47, 56, 57, 63]) for _ in range(7)]
    if scoring[0] == 3: qlens = [min(q, 40) for q in qlens]
    seqs = ["".join(rng.choice("ACGT") for _ in range(q)) for q in qlens]
    reads = []
    for n in [0, 63, 64, 65, 127, 128, 129, 191, 192, 193, 256, 320, 448, 449, 700, 1024, 1500, 2111, 3000] * mult:
        r = []
        while len(r) < n:
            s = seqs[rng.randrange(len(seqs))]
            if rng.random() < 0.3:
                s = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 40)))
            for ch in s:
                u = rng.random()
                if u < 0.04: continue
                r.append(rng.choice("ACGT") if u < 0.08 else ch)
                if u > 0.96: r.append(rng.choice("ACGT"))
        reads.append("".join(r[:n]))
    rows=[]; calls=cells=0
    for i, r in enumerate(reads):
        h, st = c_oracle.find_all(r, seqs, thr, scoring[0], scoring[1], scoring[2], scoring[2])
        h = h[np.argsort(h["start"], kind="stable")]
        calls += st["calls"]; cells += st["cells"]
        rows += [(i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"], x["score"], x["q_pos"], x["q_end"], x["depth"]) for x in h]
    want = np.array(rows, np.int32).reshape(-1, 10)
    print(qlens, len(want))
    def hm(hits):
        return np.stack([hits[k].astype(np.int32) for k in ("read","query","start","end","matches","mismatches","score","q_pos","q_end","depth")],1) if len(hits) else np.zeros((0,10),np.int32)
    if len(sys.argv) > 3:
        eng.set_scoring(scoring[0], scoring[1], scoring[2], scoring[2]); eng.set_queries(seqs); eng.upload_read_list(reads)
        import os
        for nr in ("", "1"):
            if nr: os.environ["LSW_NO_REUSE"]="1"
            else: os.environ.pop("LSW_NO_REUSE", None)
            hits = eng.find_all(thr); got = hm(hits); c = eng.last_counters
            print("gpu no_reuse=%r" % nr, got.shape, want.shape, np.array_equal(got, want), (c["tasks"], c["cells"]) == (calls, cells), c["fallbacks"])
            if not np.array_equal(got, want):
                gs = set(map(tuple, got.tolist())); ws = set(map(tuple, want.tolist()))
                print(" only gpu:", sorted(gs - ws)[:10]); print(" only oracle:", sorted(ws - gs)[:10])
    else:
        for reuse in (True, False):
            hits, cnt = emu_lib.find_all(reads, seqs, thr, *scoring, reuse=reuse)
            got = hm(hits)
            print("emu reuse", reuse, np.array_equal(got, want), (cnt["tasks"], cnt["cells"]) == (calls, cells))
            if not np.array_equal(got, want):
                gs = set(map(tuple, got.tolist())); ws = set(map(tuple, want.tolist()))
                print(" only emu:", sorted(gs - ws)[:10]); print(" only oracle:", sorted(ws - gs)[:10])
