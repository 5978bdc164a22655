"""CPU fuzzer of the chain step's device logic (lsw_chain.cuh through liblsw_emu.so): overlap pruning (lamprey.py:859-899) and
extractAmpliconAroundTarget + the stage-1 loop (lamprey.py:903-998, 1566-1589) on the hits of synthetic concatemer reads -- random
lengths, error rates 0-25 %, thresholds, overlap allowances, both shipped schemas -- against the host mirrors lamprey_b200/chain.py
and oracle/chain_ref.py.     python tools/fuzz_emu_chain.py [seconds [first seed]]"""
import ctypes
import random
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import emu_lib                                        # noqa: E402
from conftest import schema_path                      # noqa: E402
from lamprey_b200 import _native, chain, seeding, synth   # noqa: E402
from oracle import chain_ref, seeding_ref             # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = emu_lib.lib()
lib.lsw_emu_chain_subreads.restype = ctypes.c_int
t_end, it, bad, n_sub, n_hits = time.time() + budget, 0, 0, 0, 0
while time.time() < t_end:
    rng = random.Random(seed0 + it)
    schema = rng.choice(["H33", "H31"])
    ps, ps_r = seeding.PrimerSet(schema_path(schema)), seeding_ref.PrimerSet(schema_path(schema))
    pairs = ps.sequences()
    names = [n for n, _ in pairs]
    index = {n: k for k, n in enumerate(names)}
    concat, offs = synth.generate(rng.randrange(1, 12), seed=seed0 + it, mean_len=rng.choice([300, 800, 2000, 5000]),
                                  error=rng.choice([0.0, 0.03, 0.08, 0.12, 0.25]), schema=schema, min_len=30, max_len=12000,
                                  adapter_p=rng.random())
    reads = synth.reads_as_strings(concat, offs)
    thr, allowed = rng.choice([0.6, 0.7, 0.75, 0.85]), rng.choice([0, 1, 4, 4, 4, 10])
    hits, _ = emu_lib.find_all(reads, [s for _, s in pairs], thr, allowed_overlap=allowed)
    hits = np.ascontiguousarray(hits)
    n_hits += len(hits)
    fo = np.array([index.get(n, -1) for n in ps.fwd_primer_order], np.int32)
    ro = np.array([index.get(n, -1) for n in ps.rev_primer_order], np.int32)
    rlen = np.array([len(r) for r in reads], np.int32)
    out = np.zeros(len(hits) + 16, _native.SUBREAD_DT)
    n = ctypes.c_int64(0)
    rc = lib.lsw_emu_chain_subreads(emu_lib._ptr(hits), ctypes.c_int64(len(hits)), ctypes.c_int64(len(reads)), emu_lib._ptr(rlen),
                                    ctypes.c_int(len(fo)), emu_lib._ptr(fo), ctypes.c_int(len(ro)), emu_lib._ptr(ro),
                                    ctypes.c_int(index["T"]), ctypes.c_int(index["Tc"]), emu_lib._ptr(out), ctypes.c_int64(len(out)),
                                    ctypes.byref(n))
    ok = rc == 0
    got, k = out[:n.value], 0
    for i in range(len(reads)):
        h = hits[hits["read"] == i]
        al = [seeding.Alignment(int(x["start"]), int(x["end"]), x["matches"] / (x["matches"] + x["mismatches"]), names[x["query"]]) for x in h]
        # overlap pruning: the device's survivor flags == the reference function on the same list
        want_al = chain.removeOverlappingPrimerAlignments(False, list(al), allowed) if al else []
        ref_al = chain_ref.remove_overlapping_primer_alignments([seeding_ref.Alignment(a.start, a.end, a.identity, a.primer_name) for a in al], allowed) if al else []
        kept = [a for a, x in zip(al, h) if x["reserved"] == 1]
        key = lambda a: (a.start, a.end, a.primer_name)      # noqa: E731
        ok = ok and [key(a) for a in kept] == [key(a) for a in want_al] == [(a.start, a.end, a.primer_name) for a in ref_al]
        # stage-1 sub-reads around every surviving target
        want = chain.target_subreads(ps, kept, len(reads[i]), i) if kept else []
        al_r = [seeding_ref.Alignment(a.start, a.end, a.identity, a.primer_name) for a in kept]
        for a, w in zip([a for a in al_r if a.primer_name in ("T", "Tc")], want):
            s, e = chain_ref.extract_amplicon_around_target(ps_r, al_r, a)
            ok = ok and (s, len(reads[i]) - 1 if e == -1 else e) == w[:2]
        mine = [(int(x["start"]), int(x["end"]), "%d_%d" % (i, x["target_idx"]), "fwd" if x["direction"] == 0 else "rev")
                for x in got[k:k + len(want)]]
        ok = ok and mine == want and all(x["read"] == i for x in got[k:k + len(want)])
        k += len(want)
    ok = ok and k == n.value
    n_sub += k
    if not ok:
        bad += 1
        print("MISMATCH seed", seed0 + it, schema, thr, allowed, [len(r) for r in reads], flush=True)
    it += 1
print("fuzz_emu_chain: %d cases (seeds %d..%d), %d hits, %d sub-reads, %d mismatching cases" % (it, seed0, seed0 + it - 1, n_hits, n_sub, bad))
sys.exit(1 if bad else 0)
