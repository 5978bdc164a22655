import sys
sys.path.insert(0, "/root/repo")
from lamprey_b200 import _native
e = _native.Engine(0)
for m in range(9):
    r = e.measure_int_peak(m, 20000)
    print(m, r)
