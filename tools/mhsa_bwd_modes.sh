#!/bin/bash
# Per-kernel durations of the attention backward kernels (ncu, clocks untouched) in the timing-experiment modes of GVK_PIPE_DBG.
# usage: tools/mhsa_bwd_modes.sh <out.txt> [B T H]
out=$1; shift
: > "$out"
for d in 0 1 2; do
  echo "== GVK_PIPE_DBG=$d" >> "$out"
  GVK_PIPE_DBG=$d ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max --clock-control none --csv -k regex:mhsa -c 40 \
    python tools/run_mhsa.py ${@:-64 1033 12} --bwd 2>/dev/null | grep -E "mhsa" | awk -F'","' '{print $5, $(NF-2), $NF}' | sort | uniq -c | sort -k2 | awk '{print}' | tail -30 >> "$out"
done
