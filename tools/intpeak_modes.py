"""Integer-pipe microbenchmarks (lamprey_b200/csrc/intpeak.cuh + lsw.cu k_intpeak_k1row): python tools/intpeak_modes.py
modes 0-8: single instructions / simple mixes, G lane-ops/s; modes 9-12: K1's inner loop (column2), reported as GCUPS."""
import sys
sys.path.insert(0, ".")
from lamprey_b200 import _native
e = _native.Engine(0)
NAMES = {0: "VIADDMNMX.S16x2", 1: "VIMNMX3.S16x2", 2: "IADD", 3: "IMAD", 4: "row mix (old model)", 5: "row mix, IMAD sub", 6: "LOP3",
         7: "VIADDMNMX.S16x2.RELU", 8: "VIMNMX.S16x2 (2 operands)", 9: "K1 inner loop as built", 10: "K1 inner loop, keys: 2 add + VIMNMX3",
         11: "K1 inner loop, keys: 2 VIADDMNMX", 12: "K1 inner loop, keys: 2 IMAD + VIMNMX3",
         13: "K1 inner loop, keys: 2 IMAD-imm + VIMNMX3", 14: "K1 inner loop, keys and subs as IMAD-imm"}
for m in range(15):
    ops, ms = e.measure_int_peak(m, 20000 if m < 9 else 4000)
    if m >= 9:
        print("%2d %-42s %9.1f GCUPS  (%.3f ms)" % (m, NAMES[m], ops * 4 / 9 / 1e9, ms))
    else:
        print("%2d %-42s %9.1f G lane-ops/s  (%.3f ms)" % (m, NAMES[m], ops / 1e9, ms))
