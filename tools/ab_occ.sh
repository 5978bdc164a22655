#!/bin/bash
# K2 time vs resident blocks per SM (LSW_K2_BLOCKS_PER_SM) for one library
for b in 2 3 4 5 6; do
  echo "== blocks/SM $b"
  LSW_K2_BLOCKS_PER_SM=$b LSW_LIB=$PWD/$1 python tools/dbg_run.py ${2:-400000} 2>&1 | grep -E "^ok" | sed -E 's/.*(.ms_score.*ms_total.: [0-9.]+).*/\1/'
done
