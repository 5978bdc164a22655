"""Per-call latency of the module-level drop-ins (one device round trip per call): python tools/latency_align.py
swalign.LocalAlignment.align(ref, query) (lamprey.py:563) and StripedSmithWaterman(primer)(seq) (lamprey.py:577-578) on a 2 kb
read, and the batched forms beside them."""
import sys
import time
sys.path.insert(0, ".")
import numpy as np
from lamprey_b200 import seeding, skbio_ssw, swalign, synth

concat, offs = synth.generate(64, seed=5, mean_len=2000)
reads = synth.reads_as_strings(concat, offs)
ps = seeding.PrimerSet("tests/golden/H3.3_set1.primers")
seqs = [s for _, s in ps.sequences()]
sw = swalign.LocalAlignment(swalign.NucleotideScoringMatrix(2, -1))
t = ps.primer_dict["T"]
for _ in range(5):
    sw.align(reads[0], t)
    skbio_ssw.StripedSmithWaterman(t)(reads[0])
n = 200
t0 = time.perf_counter()
for i in range(n):
    sw.align(reads[i % 64], seqs[i % len(seqs)])
t1 = time.perf_counter()
for i in range(n):
    skbio_ssw.StripedSmithWaterman(seqs[i % len(seqs)])(reads[i % 64])
t2 = time.perf_counter()
b = swalign.align_many(reads, seqs)
t3 = time.perf_counter()
b2 = skbio_ssw.align_many(seqs, reads)
t4 = time.perf_counter()
al = seeding.findAllPrimerAlignments(sw, reads[0], ps, 0.7, None)
t5 = time.perf_counter()
print("LocalAlignment.align (changing primer and read every call): %.0f us per call" % ((t1 - t0) / n * 1e6))
print("StripedSmithWaterman(primer)(read):                        %.0f us per call" % ((t2 - t1) / n * 1e6))
print("align_many, 64 reads x %d sequences (swalign):            %.1f ms = %.1f us per alignment" % (len(seqs), (t3 - t2) * 1e3, (t3 - t2) / len(b) * 1e6))
print("align_many, 64 reads x %d sequences (skbio):              %.1f ms = %.1f us per alignment" % (len(seqs), (t4 - t3) * 1e3, (t4 - t3) / len(b2) * 1e6))
print("findAllPrimerAlignments, one 2 kb read (46 sequences, recursion): %.1f ms" % ((t5 - t4) * 1e3))
