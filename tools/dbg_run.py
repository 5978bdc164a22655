"""Per-round timing of one find_all (second call: buffers are allocated by the first).
usage: LSW_LIB=... [LSW_K2_DEBUG=n] python tools/dbg_run.py <reads> [workload]"""
import sys, os
import numpy as np
sys.path.insert(0, ".")
from lamprey_b200 import _native, seeding, synth
n = int(sys.argv[1])
wl = sys.argv[2] if len(sys.argv) > 2 else "synth_2kb_8pct"
concat, offs = synth.generate(n, workers=8, **synth.CONFIGS[wl])
ps = seeding.PrimerSet("tests/golden/H3.3_set1.primers")
eng = _native.Engine(0)
eng.set_scoring(2, -1, -1, -1); eng.set_queries([s for _, s in synth.schema_sequences(ps, wl)])
eng.upload_reads(concat, offs)
os.environ.pop("LSW_DEBUG", None)
eng.find_all(0.7, fetch=False)
os.environ["LSW_DEBUG"] = "1"
h = eng.find_all(0.7, fetch=False)
c = eng.last_counters
print("ok", n, h, {k: c[k] for k in ("tasks", "traced", "hits", "ms_score", "ms_trace", "ms_other", "ms_total", "fallbacks")})
