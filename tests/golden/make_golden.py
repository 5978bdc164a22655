"""Generates the committed golden fixtures under tests/golden/ (run in the build
container, where /root/reference is mounted; the GPU box never sees /root/reference).

    python tests/golden/make_golden.py

Inputs  : /root/reference/data/{H33,H31,references}  (the only fixtures the reference ships)
Outputs : reads_50.txt.gz, reads_1k.txt.gz   sequence lines of the two shipped FASTQs
          *.primers, *.fa                     the two schemas and the two target references
          align_vectors.json                  per-call known answers (Python oracle)
          seed_<reads>_<schema>_<thr>.npz     per-read seeding results (all hits, stats)

The reference has no tests and its arithmetic (swalign 0.3.6) is absent, so the "golden"
outputs are produced by oracle/swalign_ref.py + oracle/seeding_ref.py (PARITY UNPINNED,
see DESIGN.md).  The 50-read sets come from the pure-Python oracle; the 1k-read sets from
the C restatement, which this script first checks against the Python oracle on the
50-read sets, hit for hit.
"""
import gzip
import json
import os
import random
import shutil
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/data"

from oracle import swalign_ref, seeding_ref, c_oracle  # noqa: E402

SCHEMAS = {"H33": "H33/H3.3_set1.primers", "H31": "H31/H3.1_set1.primers"}
THRS = (0.70, 0.75)


def fastq_seqs(path):
    seqs = []
    with open(path) as fp:
        for i, line in enumerate(fp):
            if i % 4 == 1:
                seqs.append(line.strip())
    return seqs


def py_seed(args):
    seq, schema_path, thr = args
    aligner = swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1))
    ps = seeding_ref.PrimerSet(schema_path)
    names = [n for n, _ in ps.sequences()]
    st = seeding_ref.Stats()
    hits, cov = seeding_ref.find_all_primer_alignments(aligner, seq, ps, thr, st)
    return [(names.index(h.primer_name), h.start, h.end, h.identity) for h in hits], cov, st.calls, st.cells


def c_seed(seq, seqs, thr):
    hits, st = c_oracle.find_all(seq, seqs, thr)
    order = np.argsort(hits["start"], kind="stable")     # lamprey.py:722 stable sort on start
    return hits[order], st


def main():
    for rel in list(SCHEMAS.values()) + ["references/H3F3A.fa", "references/HIST1H3B.fa"]:
        shutil.copyfile(os.path.join(REF, rel), os.path.join(HERE, os.path.basename(rel)))
    reads = {"50": fastq_seqs(os.path.join(REF, "H33/LAMP_14min_50.fastq")),
             "1k": fastq_seqs(os.path.join(REF, "H33/LAMP_14min_1k.fastq"))}
    for k, v in reads.items():
        with gzip.open(os.path.join(HERE, "reads_%s.txt.gz" % k), "wt") as fp:
            fp.write("\n".join(v) + "\n")

    # ---- per-call vectors ---------------------------------------------------------------
    al = swalign_ref.LocalAlignment(swalign_ref.NucleotideScoringMatrix(2, -1))
    rng = random.Random(20261019)
    cases = [("ACACACTA", "AGCACACA"),            # upstream README / lamprey.py:1860-1865
             ("AAAA", "CCCC"), ("", "ACGT"), ("ACGT", ""), ("", ""),
             ("ACGTACGTAC", "ACGTACGTAC"), ("acgtNNacgt", "ACGTNNACGT"), ("ACGTNACGT", "ACGTAACGT"),
             ("A" * 40, "A" * 12), ("AC" * 30, "CA" * 8), ("ACGT" * 10, "ACGGT")]
    ps = seeding_ref.PrimerSet(os.path.join(REF, SCHEMAS["H33"]))
    seqs33 = ps.sequences()
    for i in range(150):
        read = reads["50"][rng.randrange(50)]
        a = rng.randrange(0, max(1, len(read) - 80))
        b = min(len(read), a + rng.randrange(30, 400))
        cases.append((read[a:b], seqs33[rng.randrange(len(seqs33))][1]))
    for i in range(150):
        alpha = rng.choice(["AC", "A", "ACG", "ACGT", "ACGTN"])
        ref = "".join(rng.choice(alpha) for _ in range(rng.randrange(1, 160)))
        q = "".join(rng.choice(alpha) for _ in range(rng.randrange(1, 40)))
        cases.append((ref, q))
    vecs = []
    for ref, q in cases:
        a = al.align(ref, q)
        c = c_oracle.align(ref, q)
        assert (a.score, a.q_pos, a.q_end, a.r_pos, a.r_end, a.matches, a.mismatches, a.cigar) == \
               (c["score"], c["q_pos"], c["q_end"], c["r_pos"], c["r_end"], c["matches"], c["mismatches"], c["cigar"])
        vecs.append(dict(ref=ref, query=q, score=a.score, q_pos=a.q_pos, q_end=a.q_end, r_pos=a.r_pos,
                         r_end=a.r_end, matches=a.matches, mismatches=a.mismatches, cigar=a.cigar_str,
                         identity=a.identity))
    with open(os.path.join(HERE, "align_vectors.json"), "w") as fp:
        json.dump(vecs, fp, indent=0)

    # ---- seeding goldens ----------------------------------------------------------------
    pool = Pool()
    for sname, rel in SCHEMAS.items():
        spath = os.path.join(REF, rel)
        ps = seeding_ref.PrimerSet(spath)
        names = [n for n, _ in ps.sequences()]
        seqs = [s for _, s in ps.sequences()]
        for thr in THRS:
            py = pool.map(py_seed, [(r, spath, thr) for r in reads["50"]])
            for rk, rds in reads.items():
                rows, cov, calls, cells = [], [], 0, 0
                for i, r in enumerate(rds):
                    h, st = c_seed(r, seqs, thr)
                    tot = h["matches"] + h["mismatches"]
                    ident = h["matches"] / np.maximum(tot, 1)
                    if rk == "50":      # C restatement vs Python oracle, hit for hit
                        ph, pcov, pcalls, pcells = py[i]
                        assert [(int(x["seq"]), int(x["start"]), int(x["end"])) for x in h] == \
                               [(a, b, c) for a, b, c, _ in ph], (sname, thr, i)
                        assert [float(v) for v in ident] == [d for _, _, _, d in ph]
                        assert (pcalls, pcells) == (st["calls"], st["cells"])
                    for x in h:
                        rows.append((i, x["seq"], x["start"], x["end"], x["matches"], x["mismatches"],
                                     x["score"], x["q_pos"], x["q_end"], x["depth"]))
                    cov.append(float((h["end"] - h["start"]).sum()) / float(len(r)))
                    calls += st["calls"]
                    cells += st["cells"]
                arr = np.array(rows, dtype=np.int32).reshape(-1, 10)
                out = os.path.join(HERE, "seed_%s_%s_%02d.npz" % (rk, sname, round(thr * 100)))
                np.savez_compressed(out, hits=arr, coverage=np.array(cov), calls=calls, cells=cells,
                                    names=np.array(names), thr=thr,
                                    columns=np.array(["read", "seq", "start", "end", "matches", "mismatches",
                                                      "score", "q_pos", "q_end", "depth"]))
                print(out, len(arr), calls, cells, flush=True)


if __name__ == "__main__":
    main()
