#!/bin/bash
# per-round K1/K2 times of several library builds on the same box: tools/ab_rounds.sh <reads> "<debug modes>" lib1.so lib2.so ...
n=$1; modes=$2; shift; shift
for lib in "$@"; do
  for d in $modes; do
    echo "== $lib LSW_K2_DEBUG=$d"
    LSW_LIB=$PWD/$lib LSW_K2_DEBUG=$d python tools/dbg_run.py $n 2>&1 | grep -E "round [0-9]+:|^ok" | sed -E 's/.*round ([0-9]+): tasks ([0-9]+).*score ([0-9.]+) ms, trace ([0-9.]+) ms, other ([0-9.]+) ms.*/r\1 tasks \2 score \3 trace \4 other \5/' | head -${ROUNDS:-4}
  done
done
