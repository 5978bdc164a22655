"""One find_all on the skbio path (for ncu captures and timing): python tools/dbg_ssw.py <reads> [scalar]"""
import os
import sys
import time
sys.path.insert(0, ".")
from lamprey_b200 import _native, seeding, synth
n = int(sys.argv[1])
if len(sys.argv) > 2 and sys.argv[2] == "scalar":
    os.environ["LSW_SSW_SCALAR"] = "1"
concat, offs = synth.generate(n, workers=8, **synth.CONFIGS["synth_2kb_8pct"])
ps = seeding.PrimerSet("tests/golden/H3.3_set1.primers")
eng = _native.Engine(0)
eng.set_ssw(2, -3, 5, 2); eng.set_queries([s for _, s in ps.sequences()]); eng.upload_reads(concat, offs)
eng.find_all(0.7, fetch=False)
t0 = time.perf_counter(); h = eng.find_all(0.7, fetch=False); dt = time.perf_counter() - t0
c = eng.last_counters
print("ok skbio path", n, "reads:", h, "hits,", c["tasks"], "tasks, %.1f ms, %.1f GCUPS, %d rounds" % (dt * 1e3, c["cells"] / dt / 1e9, c["rounds"]))
