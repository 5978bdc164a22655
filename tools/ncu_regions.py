"""Summarise `ncu --page source --csv` of one kernel by SASS regions: python tools/ncu_regions.py file.csv [n_regions_by_backward_branches]
Prints, per region between loop boundaries (backward branches), samples, executed instructions, average active threads, top stalls."""
import csv, sys, re
allrows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0          # which kernel of the report
rows = allrows[starts[which]:starts[which + 1]]
print(rows[0][1])
h = rows[1]
col = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
addr = [int(r[col["Address"]], 16) for r in data]
base = addr[0]
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
def num(r, k):
    try: return float(r[col[k]])
    except Exception: return 0.0
# region boundaries: targets of backward branches and the branches themselves
bounds = set([0, len(data)])
for i, r in enumerate(data):
    m = re.search(r"BRA\s+(?:\S+,\s*)?0x([0-9a-f]+)", r[col["Source"]])
    if m:
        t = int(m.group(1), 16)
        if t > base: t -= base           # ncu prints absolute targets
        if t <= addr[i] - base:            # backward
            j = next((k for k, a in enumerate(addr) if a - base == t), None)
            if j is not None: bounds.add(j); bounds.add(i + 1)
b = sorted(bounds)
tot = sum(num(r, "# Samples") for r in data)
print("total samples %d, instructions executed %.0f M" % (tot, sum(num(r, "Instructions Executed") for r in data) / 1e6))
for lo, hi in zip(b[:-1], b[1:]):
    seg = data[lo:hi]
    s = sum(num(r, "# Samples") for r in seg)
    if s < 0.005 * tot: continue
    ie = sum(num(r, "Instructions Executed") for r in seg); te = sum(num(r, "Thread Instructions Executed") for r in seg)
    st = sorted(((sum(num(r, k) for r in seg), k) for k in stalls), reverse=True)[:4]
    print("[%5d,%5d) 0x%04x-0x%04x  %5.1f %% samples  %7.1f M inst  %4.1f thr/inst  %s" % (
        lo, hi, addr[lo] - base, addr[hi - 1] - base, 100 * s / tot, ie / 1e6, te / max(ie, 1), ", ".join("%s %d" % (k[6:], v) for v, k in st)))
