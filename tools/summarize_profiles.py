#!/usr/bin/env python
"""Turns the ncu captures that tools/profile_cmd.sh leaves in gpurun_out/ into the tracked summaries
under profiles/ (run here, on the build box: `ncu -i` needs no GPU).

    python tools/summarize_profiles.py r01

writes profiles/<tag>_launches.csv (every launch with its device time), profiles/<tag>_kernels.md
(share of the step per kernel + the full-set metrics of the hot kernels) and keeps the raw CSV of
the metrics it quotes in profiles/<tag>_<kernel>_raw.csv.
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

DETAILS = ["Duration", "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Issue Slots Busy",
           "Executed Ipc Active", "No Eligible", "Avg. Active Threads Per Warp", "L1/TEX Hit Rate", "L2 Hit Rate",
           "DRAM Throughput", "Compute (SM) Throughput", "Grid Size", "Dynamic Shared Memory Per Block"]
RAW = ["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "gpu__time_duration.sum",
       "sm__cycles_elapsed.avg", "smsp__inst_executed.sum"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(tag, md, fname="launches.csv", suffix="", cmd="python bench.py --reads 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e", first=400):
    path = os.path.join(GP, fname)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[h], rows[h + 1:]
    ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
    tot, cnt, seq = defaultdict(float), defaultdict(int), []
    for r in data:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v
        name = r[ik].split("(")[0].replace("void ", "").replace("lsw::", "").replace("<unnamed>::", "")
        if name.startswith("cub::"):
            name = name.split("<")[0] + ("<u16 keys>" if "unsigned short" in r[ik] else "<u64 keys>" if "unsigned long" in r[ik] else "")
        tot[name] += v
        cnt[name] += 1
        seq.append((r[0], name, v))
    with open(os.path.join(OUT, tag + "_launches%s.csv" % suffix), "w") as fp:
        fp.write("id,kernel,duration_us\n")
        for i, n, v in seq:
            fp.write("%s,%s,%.3f\n" % (i, n, v))
    step = {k: v for k, v in tot.items() if not k.startswith("k_intpeak")}
    total = sum(step.values())
    md.append("## Launch list%s (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised)\n" % suffix.replace("_", " "))
    md.append("Command: `%s` (first %d launches: roofline microbenchmark, warm-up step, timed step). Shares are of the "
              "launches that belong to a step (the `k_intpeak` roofline microbenchmark is excluded).\n" % (cmd, first))
    md.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(step.items(), key=lambda x: -x[1]):
        md.append("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], v / 1e3, 100 * v / total))
    md.append("")


def kernel(tag, name, md):
    rep = os.path.join(GP, "prof_%s.ncu-rep" % name)
    if not os.path.exists(rep):
        return
    det = ncu_csv(rep, "details")
    hdr = det[0]
    ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    per = defaultdict(dict)
    names = {}
    for r in det[1:]:
        if r[mi] in DETAILS:
            per[r[ii]][r[mi]] = r[vi] + " " + r[ui]
            names[r[ii]] = r[ki].split("(")[0].replace("void ", "")
    raw = ncu_csv(rep, "raw")
    rh = raw[0]
    with open(os.path.join(OUT, "%s_%s_raw.csv" % (tag, name)), "w") as fp:
        w = csv.writer(fp)
        cols = [rh.index("ID"), rh.index("Kernel Name")] + [rh.index(k) for k in RAW if k in rh]
        for r in raw:
            w.writerow([r[c] for c in cols])
    md.append("## `%s` (`ncu --set full --clock-control none --import-source on`, 3 launches)\n" % name)
    keys = list(DETAILS)
    md.append("| metric | " + " | ".join("`%s`" % names[i] for i in sorted(per)) + " |")
    md.append("|---|" + "---|" * len(per))
    for k in keys:
        md.append("| %s | " % k + " | ".join(per[i].get(k, "") for i in sorted(per)) + " |")
    for k in RAW:
        if k in rh:
            c = rh.index(k)
            md.append("| `%s` | " % k + " | ".join(r[c] for r in raw[2:]) + " |")
    md.append("")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    md = ["# %s - ncu summaries (B200, sm_100a)\n" % tag,
          "Produced by `tools/profile_cmd.sh` under gpurun and `tools/summarize_profiles.py` here. "
          "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n"]
    launches(tag, md)
    launches(tag, md, "launches_1M.csv", "_1M", "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e", 1200)
    kernel(tag, "score", md)
    kernel(tag, "trace", md)
    with open(os.path.join(OUT, tag + "_kernels.md"), "w") as fp:
        fp.write("\n".join(md) + "\n")
    print("\n".join(md)[:3000])


if __name__ == "__main__":
    main()
