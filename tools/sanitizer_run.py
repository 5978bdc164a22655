"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): the 50-read fixture through find_all (swalign and skbio
paths, overlap pruning, device sub-reads) and stage 2.   compute-sanitizer --tool racecheck python tools/sanitizer_run.py"""
import gzip
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lamprey_b200 import pipeline, seeding  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
reads = gzip.open(os.path.join(GOLD, "reads_50.txt.gz"), "rt").read().split()[:int(sys.argv[1]) if len(sys.argv) > 1 else 50]
ps = seeding.PrimerSet(os.path.join(GOLD, "H3.3_set1.primers"))
ref = open(os.path.join(GOLD, "H3F3A.fa")).read().split("\n", 1)[1].replace("\n", "").upper()
seeds = pipeline.seed_batch(reads, ps, 0.70)
recs = pipeline.stage2_align_batch(reads, seeds, ref)
sk = seeding.seed_reads(reads, ps, 0.70, method="skbio")
print("ok: %d hits, %d sub-reads, %d stage-2 alignments, %d skbio-path hits" % (len(seeds.hits), len(seeds.subreads), len(recs), len(sk.hits)))
