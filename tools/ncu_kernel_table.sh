#!/bin/bash
# Per-kernel table of GAViKO training steps (B200, ncu): duration, DRAM / SM throughput, occupancy, registers, tensor pipe.
# usage: tools/ncu_kernel_table.sh <out.csv> [one_step.py args]     (profiles the 2nd of 2 steps)
out=$1; shift
python tools/one_step.py "$@" > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file "$out" python tools/one_step.py "$@" > gpurun_out/ncu_run.log 2>&1
