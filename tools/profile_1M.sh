#!/bin/bash
# ncu launch list + full captures of the hot kernels AT THE 10^6-READ CONFIG (run under gpurun): tools/profile_1M.sh <tag>
# gpurun copies back at most 64 MiB: the reports are turned into CSV pages on the box and dropped if they are too large.
tag=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --no-strong"
O=gpurun_out
if [ -z "$NO_LAUNCH_LIST" ]; then
$CMD > $O/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${tag}_launches_1M.csv $CMD > $O/${tag}_ncu_launches.log 2>&1
fi
cap() {   # cap <name> <kernel base name> <launches of that kernel to skip>
  $CMD > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k $2 --launch-skip $3 -c 1 -o $O/${tag}_prof_$1 $CMD > $O/${tag}_ncu_$1.log 2>&1
  ncu -i $O/${tag}_prof_$1.ncu-rep --page raw --csv > $O/${tag}_$1_raw.csv 2>/dev/null
  ncu -i $O/${tag}_prof_$1.ncu-rep --page details --csv > $O/${tag}_$1_details.csv 2>/dev/null
  ncu -i $O/${tag}_prof_$1.ncu-rep --page source --csv > $O/${tag}_$1_source.csv 2>/dev/null
  gzip -f $O/${tag}_$1_source.csv
  sz=$(stat -c %s $O/${tag}_prof_$1.ncu-rep 2>/dev/null || echo 0); [ "$sz" -gt 12000000 ] && rm -f $O/${tag}_prof_$1.ncu-rep
}
# round 0 of the warm-up step: the first launch of the 24-row class (whole reads), K1 and K2
# (kernels of a round are launched in class order 16, 20, 24, 56, 64: skip two launches of the same base name)
cap score24 k_score 2
cap band24 k_trace_band 2
cap band16 k_trace_band 0
# the skbio path: round 0 of the packed affine K1, 24-row class, 10^5 reads
CMD="python tools/dbg_ssw.py 100000"
cap affine24 k_score_affine 2
for f in $O/${tag}_ncu_*.log; do echo "== $f"; tail -n 3 $f; done
du -sh $O; ls -la $O
