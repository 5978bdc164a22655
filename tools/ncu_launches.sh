#!/bin/bash
# ncu launch list of one find_all (second call of tools/dbg_run.py): tools/ncu_launches.sh <reads> <out-prefix> [workload]
n=${1:-200000}; out=${2:-gpurun_out/launches}; wl=${3:-synth_2kb_8pct}
python tools/dbg_run.py $n $wl > ${out}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${out}.csv python tools/dbg_run.py $n $wl > ${out}_ncu.log 2>&1
python - <<PY
import csv, collections, re
rows = list(csv.reader(l for l in open("${out}.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("lsw::", "")
    name = re.sub(r"cub::CUB_[0-9_A-Za-z]*::", "cub::", name)[:70]
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-72s %5d %10.3f ms %5.1f %%" % (k, a[0], a[1], 100 * a[1] / tot))
print("total %.3f ms over %d launches (both find_all calls of dbg_run)" % (tot, sum(a[0] for a in agg.values())))
PY
