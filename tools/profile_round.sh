#!/bin/bash
# Round profile pass (run under gpurun): launch list of the default bench command, GEMM DRAM traffic, one `--set full` capture per hot kernel.
# usage: tools/profile_round.sh r02
tag=$1
mkdir -p gpurun_out
B="--steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline"
python bench.py $B > gpurun_out/bench_plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_${tag}_bench_launches.csv python bench.py $B > gpurun_out/ncu_bench_run.log 2>&1
python tools/one_step.py --batch 64 > gpurun_out/one_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/gemm_traffic_$tag.csv \
    python tools/one_step.py --batch 64 > gpurun_out/ncu_gemm_run.log 2>&1
# one full capture per hot kernel of the second step (-s skips the first step's launches of that kernel)
for spec in "mhsa_bwd_dq_pipe:12" "mhsa_bwd_dkv_pipe:12" "mhsa_ws_fwd:12" "patch_embed_tf32:1" "layernorm_bwd:36" "tc_wgrad:60" "tc_down:72" "win_dkv:12"; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/full_${tag}_$k python tools/one_step.py --batch 64 > gpurun_out/ncu_full_$k.log 2>&1
done
# the two GELU-epilogue GEMMs and a plain one: launch order inside a layer's forward is qkv, out-proj, fc1 (GELU + saved derivative), fc2
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 98 -c 4 -o gpurun_out/full_${tag}_gemm_fwd_layer0 python tools/one_step.py --batch 64 > gpurun_out/ncu_full_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
