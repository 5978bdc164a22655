"""profiles/r02_kernels.md from the round-2 ncu captures at the 10^6-read config (tools/profile_1M.sh) -- run here, no GPU needed.
    python tools/summarize_r02.py [tag under gpurun_out/, default: files already copied to profiles/r02_*]"""
import collections
import csv
import gzip
import io
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
DETAILS = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Issue Slots Busy", "Executed Ipc Active",
           "No Eligible", "Avg. Active Threads Per Warp", "L1/TEX Hit Rate", "L2 Hit Rate", "DRAM Throughput", "Compute (SM) Throughput",
           "Grid Size", "Dynamic Shared Memory Per Block", "Block Limit Registers", "Block Limit Shared Mem"]
RAW = ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "dram__bytes_read.sum", "dram__bytes_write.sum",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
md = ["# r02 - ncu summaries (B200, sm_100a) at the 10^6-read configuration\n",
      "Produced by `tools/profile_1M.sh` under gpurun (`python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-strong`) and "
      "`tools/summarize_r02.py` here. Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n"]

# ---- launch list ------------------------------------------------------------------------------------------------------------
path = os.path.join(P, "r02_launches_1M_ncu.csv")
if os.path.exists(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("lsw::", "").replace("<unnamed>::", "")
        name = re.sub(r"cub::CUB_[0-9_A-Za-z]*::", "cub::", name)
        if name.startswith("cub::"):
            name = name.split("<")[0] + ("<u16 keys>" if "unsigned short" in r[ki] else "<u64 keys>" if "unsigned long" in r[ki] else "")
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        seq.append((name, ms))
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_init_tasks")]
    step = seq[starts[-1]:] if starts else seq
    agg = collections.OrderedDict()
    for n, ms in step:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    md.append("## Launch list of one step (`ncu --metrics gpu__time_duration.sum --clock-control none`): %d launches, %.1f ms\n" % (len(step), tot))
    md.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if a[1] / tot >= 0.0005:
            md.append("| `%s` | %d | %.3f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
    md.append("")
    rnd, acc = 0, {}
    for n, ms in step:
        key = "K1" if "k_score" in n else "K2" if "k_trace" in n else "other"
        acc.setdefault(rnd, {"K1": 0.0, "K2": 0.0, "other": 0.0})[key] += ms
        if "k_scatter_children" in n:
            rnd += 1
    md.append("Per recursion round (ms): " + "; ".join("r%d K1 %.1f K2 %.1f other %.1f" % (r, v["K1"], v["K2"], v["other"]) for r, v in sorted(acc.items()) if sum(v.values()) > 0.5) + "\n")

# ---- full captures ----------------------------------------------------------------------------------------------------------
for key, title in (("score24", "k_score<24>, round 0 (32 sequences x 10^6 whole reads)"), ("band24", "k_trace_band<24,2,0>, round 0"),
                   ("band16", "k_trace_band<16,2,0>, round 0"), ("affine24", "k_score_affine<24>, round 0 (skbio path, 10^5 reads)")):
    dpath, rpath = os.path.join(P, "r02_%s_1M_details.csv" % key), os.path.join(P, "r02_%s_1M_raw.csv" % key)
    if not (os.path.exists(dpath) and os.path.exists(rpath)):
        continue
    det = list(csv.reader(open(dpath)))
    h = det[0]
    mi, vi, ui = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    md.append("## `%s` (`ncu --set full --clock-control none --import-source on`)\n" % title)
    md.append("| metric | value |\n|---|---|")
    seen = set()
    for r in det[1:]:
        if r[mi] in DETAILS and r[mi] not in seen:
            seen.add(r[mi]); md.append("| %s | %s %s |" % (r[mi], r[vi], r[ui]))
    raw = list(csv.reader(open(rpath)))
    for k in RAW:
        if k in raw[0]:
            c = raw[0].index(k); md.append("| `%s` | %s %s |" % (k, raw[2][c], raw[1][c]))
    spath = os.path.join(P, "r02_%s_1M_source.csv.gz" % key)
    if os.path.exists(spath):
        rows = list(csv.reader(io.TextIOWrapper(gzip.open(spath))))
        st = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
        hh = rows[st[0] + 1]
        data = [r for r in rows[st[0] + 2:] if len(r) == len(hh)]
        col = {k: i for i, k in enumerate(hh)}
        mix = collections.Counter()
        for r in data:
            parts = r[col["Source"]].split()
            op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
            try:
                mix[op] += float(r[col["Instructions Executed"]])
            except ValueError:
                pass
        t = sum(mix.values())
        md.append("\nExecuted instruction mix (ncu source page): " + ", ".join("%s %.1f %%" % (k, 100 * v / t) for k, v in mix.most_common(9)) + "\n")
    md.append("")
with open(os.path.join(P, "r02_kernels.md"), "w") as fp:
    fp.write("\n".join(md) + "\n")
print("\n".join(md)[:2500])
