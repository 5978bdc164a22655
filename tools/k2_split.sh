#!/bin/bash
# K2 time split, round 0 only is comparable (debug modes emit no hits): normal / fill only / fill without stores
for d in 0 1 2; do
  echo "== LSW_K2_DEBUG=$d"
  LSW_K2_DEBUG=$d LSW_DEBUG=1 python tools/dbg_run.py ${1:-400000} 2>&1 | grep -E "round 0:|^ok" | cut -c1-200
done
