#!/bin/bash
# ncu launch list + full captures of the hot kernels AT THE 10^6-READ CONFIG (run under gpurun): tools/profile_1M.sh <tag>
tag=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_1M.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
# round 0 of the warm-up step: the first k_score launch of every register-row class (whole reads)
ncu --set full --clock-control none --import-source on -k regex:k_score -c 5 -o gpurun_out/${tag}_prof_score_1M $CMD > gpurun_out/${tag}_ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_band -c 5 -o gpurun_out/${tag}_prof_band_1M $CMD > gpurun_out/${tag}_ncu_band.log 2>&1
tail -3 gpurun_out/${tag}_ncu_launches.log gpurun_out/${tag}_ncu_score.log gpurun_out/${tag}_ncu_band.log
ls -la gpurun_out/
