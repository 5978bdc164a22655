"""Round-2 additions to the committed goldens (needs neither /root/reference nor a GPU: inputs are the committed fixtures and
the seeded synthetic generator).

    python tests/golden/make_golden_more.py

  ssw_vectors.json                        per-call known answers of the skbio / SSW restatement (oracle/ssw_oracle.c), first of
                                          them the example of scikit-bio's own documentation
  seed_<reads>_<schema>_<thr>_skbio.npz   per-read seeding results of the DEFAULT (skbio) path, lamprey.py:596-645 with
                                          args.swalign False
  synth_digests.json                      swalign-path seeding of 1000-read slices of BASELINE.json configs[3] and [4]
                                          (5 kb / 12 %; 20-50 kb, 40-sequence schema): hit count, align() calls, cells and the
                                          SHA-256 of the int32 hit matrix (the matrices themselves are tens of MB)
All produced by the oracle (PARITY UNPINNED)."""
import gzip
import hashlib
import json
import os
import random
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import c_oracle, ssw_oracle, seeding_ref  # noqa: E402
from lamprey_b200 import synth  # noqa: E402   (the seeded read generator; no alignment code)

DOC_TARGET = "AAAAAACTCTCTAAACTCACTAAGGCTCTCTACCCCTCTTCAGAGAAGTCGA"
DOC_QUERY = "ACTAAGGCTCTCTACCCCTCTCAGAGA"


def load_reads(key):
    with gzip.open(os.path.join(HERE, "reads_%s.txt.gz" % key), "rt") as fp:
        return fp.read().split()


def ssw_seed(args):
    read, seqs, thr = args
    h, st = ssw_oracle.find_all(read, seqs, thr)
    return h[np.argsort(h["start"], kind="stable")], st


def sw_seed(args):
    read, seqs, thr, idx = args
    h, st = c_oracle.find_all(read, seqs, thr)
    h = h[np.argsort(h["start"], kind="stable")]
    rows = np.zeros((len(h), 10), np.int32)
    rows[:, 0] = idx
    for k, name in enumerate(("seq", "start", "end", "matches", "mismatches", "score", "q_pos", "q_end", "depth")):
        rows[:, k + 1] = h[name]
    return rows, st["calls"], st["cells"]


def main():
    c_oracle.build()
    ssw_oracle.build()
    rng = random.Random(20261020)
    # ---- per-call SSW vectors -----------------------------------------------------------------
    cases = [(DOC_TARGET, DOC_QUERY), ("AAAA", "CCCC"), ("ACGTACGTAC", "ACGTACGTAC"), ("ACGTNNACGT", "ACGTACGT"),
             ("A" * 40, "A" * 12), ("AC" * 30, "CA" * 8), ("ACGT" * 10, "ACGGT"), ("T", "T"), ("ACGT", "T")]
    reads50 = load_reads("50")
    ps33 = seeding_ref.PrimerSet(os.path.join(HERE, "H3.3_set1.primers"))
    seqs33 = [s for _, s in ps33.sequences()]
    for _ in range(150):
        read = reads50[rng.randrange(50)]
        a = rng.randrange(0, max(1, len(read) - 80))
        b = min(len(read), a + rng.randrange(30, 400))
        cases.append((read[a:b], seqs33[rng.randrange(len(seqs33))]))
    for _ in range(150):
        alpha = rng.choice(["AC", "A", "ACG", "ACGT", "ACGTN"])
        t = "".join(rng.choice(alpha) for _ in range(rng.randrange(1, 160)))
        q = "".join(rng.choice(alpha.replace("N", "") or "A") for _ in range(rng.randrange(1, 40)))
        cases.append((t, q))
    vecs = []
    for t, q in cases:
        d = ssw_oracle.align(t, q)
        vecs.append(dict(target=t, query=q, score=d["score"], target_begin=d["ref_begin"], target_end_optimal=d["ref_end"],
                         query_begin=d["read_begin"], query_end=d["read_end"], suboptimal_score=d["score2"],
                         target_end_suboptimal=d["ref_end2"], matches=d["matches"],
                         cigar="".join("%d%s" % x for x in d["cigar"])))
    with open(os.path.join(HERE, "ssw_vectors.json"), "w") as fp:
        json.dump(vecs, fp, indent=0)
    pool = Pool()
    # ---- default-path seeding goldens ---------------------------------------------------------
    for sname, fn in (("H33", "H3.3_set1.primers"), ("H31", "H3.1_set1.primers")):
        ps = seeding_ref.PrimerSet(os.path.join(HERE, fn))
        names = [n for n, _ in ps.sequences()]
        seqs = [s for _, s in ps.sequences()]
        for rk in ("50", "1k"):
            rds = load_reads(rk)
            for thr in (0.70, 0.75):
                res = pool.map(ssw_seed, [(r, seqs, thr) for r in rds])
                rows, cov, calls, cells = [], [], 0, 0
                for i, (h, st) in enumerate(res):
                    for x in h:
                        rows.append((i, x["seq"], x["start"], x["end"], x["matches"], x["score"], x["q_pos"], x["q_end"], x["depth"]))
                    cov.append(float((h["end"] - h["start"]).sum()) / float(len(rds[i])))
                    calls += st["calls"]
                    cells += st["cells"]
                arr = np.array(rows, dtype=np.int32).reshape(-1, 9)
                out = os.path.join(HERE, "seed_%s_%s_%02d_skbio.npz" % (rk, sname, round(thr * 100)))
                np.savez_compressed(out, hits=arr, coverage=np.array(cov), calls=calls, cells=cells, names=np.array(names), thr=thr,
                                    columns=np.array(["read", "seq", "start", "end", "matches", "score", "q_pos", "q_end", "depth"]))
                print(out, len(arr), calls, cells, flush=True)
    # ---- 1000-read slices of the two large synthetic configs (swalign path) --------------------
    digests = {}
    for wl in ("synth_5kb_12pct", "synth_20_50kb_12pct"):
        n = 1000
        concat, offs = synth.generate(n, **synth.CONFIGS[wl])
        reads = synth.reads_as_strings(concat, offs)
        seqs = [s for _, s in synth.schema_sequences(ps33, wl)]
        res = pool.map(sw_seed, [(r, seqs, 0.7, i) for i, r in enumerate(reads)], chunksize=4)
        mat = np.concatenate([r[0] for r in res]) if res else np.zeros((0, 10), np.int32)
        digests[wl] = {"reads": n, "threshold": 0.7, "sequences": len(seqs), "bases": int(offs[-1]), "hits": int(len(mat)),
                       "calls": int(sum(r[1] for r in res)), "cells": int(sum(r[2] for r in res)),
                       "sha256": hashlib.sha256(np.ascontiguousarray(mat, dtype="<i4").tobytes()).hexdigest(),
                       "columns": ["read", "query", "start", "end", "matches", "mismatches", "score", "q_pos", "q_end", "depth"],
                       "first_rows": mat[:8].tolist()}
        print(wl, digests[wl]["hits"], digests[wl]["calls"], digests[wl]["cells"], flush=True)
    with open(os.path.join(HERE, "synth_digests.json"), "w") as fp:
        json.dump(digests, fp, indent=1)


if __name__ == "__main__":
    main()
