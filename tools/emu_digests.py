"""The host emulation of the kernels (liblsw_emu.so) against the committed oracle digests of the 1000-read slices of BASELINE configs 4 and 5
(tests/golden/synth_digests.json; the GPU runs the same comparison in tests/test_gpu_parity.py).  ~2 minutes of CPU.   python tools/emu_digests.py"""
import sys, time, json, hashlib; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, emu_lib
from conftest import hits_matrix
from lamprey_b200 import synth, seeding
ps=seeding.PrimerSet("tests/golden/H3.3_set1.primers")
gold_all=json.load(open("tests/golden/synth_digests.json"))
for wl,gold in gold_all.items():
    c,o=synth.generate(gold["reads"], workers=4, **synth.CONFIGS[wl])
    assert int(o[-1])==gold["bases"]
    reads=synth.reads_as_strings(c,o)
    seqs=[s for _,s in synth.schema_sequences(ps,wl)]
    t=time.time(); hits,cnt=emu_lib.find_all(reads,seqs,gold["threshold"],cap=gold["hits"]+16); dt=time.time()-t
    mat=hits_matrix(hits)
    ok=(len(mat),cnt["tasks"],cnt["cells"])==(gold["hits"],gold["calls"],gold["cells"]) and hashlib.sha256(np.ascontiguousarray(mat,dtype="<i4").tobytes()).hexdigest()==gold["sha256"]
    print(wl, "emulation == committed oracle digest:", ok, "hits", len(mat), "fallbacks", cnt["fallbacks"], "computed/ref %.3f"%(cnt["computed_cells"]/cnt["cells"]), "%.0fs"%dt, flush=True)
