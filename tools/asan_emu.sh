#!/bin/bash
# Bounds-check the kernels' lane logic on the CPU: the __host__ __device__ code of K1/K2 (lsw_score.cuh,
# lsw_trace.cuh) built into the host emulation with AddressSanitizer.  (compute-sanitizer is closed on the
# GPU pool, so index arithmetic is checked here.)  Run from the repo root.
set -e
OUT=${TMPDIR:-/tmp}/liblsw_emu_asan.so
( cd lamprey_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 \
    -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -shared -diag-suppress 20013,20015 \
    -o "$OUT" lsw_emu.cu -Xlinker -lasan )
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 LSW_EMU_LIB="$OUT" python - <<'PY'
import gzip, os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import emu_lib
emu_lib.PATH = os.environ["LSW_EMU_LIB"]
from oracle import seeding_ref
from lamprey_b200 import synth
from lamprey_b200._native import TASK_DT
G = "tests/golden/"
seqs = [s for _, s in seeding_ref.PrimerSet(G + "H3.3_set1.primers").sequences()]
reads = gzip.open(G + "reads_50.txt.gz", "rt").read().split()
print(emu_lib.find_all(reads, seqs, 0.7, allowed_overlap=4)[1])
c, o = synth.generate(6, **synth.CONFIGS["synth_2kb_8pct"])
print(emu_lib.find_all(synth.reads_as_strings(c, o), seqs, 0.7)[1])
print(emu_lib.find_all(["", "A", "ACGTACGTACGTACGTACGT", seqs[0] * 3, "N" * 70], seqs, 0.7, force_full=True)[1])
t = np.array([(0, 0, len(reads[0]), 3), (1, 5, 60, 7), (2, 0, 0, 1)], dtype=TASK_DT)
print(emu_lib.align_tasks(reads, seqs, t)[0]["score"])
print("ASAN: clean")
PY
