#!/bin/bash
# ncu captures of the two hot kernels (run under gpurun; see /opt/skills/guides/B200_PROFILING.md)
set -x
CMD="python bench.py --reads 32768 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_score -c 3 -o gpurun_out/prof_score $CMD > gpurun_out/ncu_score.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -c 3 -o gpurun_out/prof_trace $CMD > gpurun_out/ncu_trace.log 2>&1
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_score.log gpurun_out/ncu_trace.log
ls -la gpurun_out/
