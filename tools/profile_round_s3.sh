#!/bin/bash
# Full ncu captures of the kernels written / rewritten in the last session of round 2 (run under gpurun): EVP's two primitives inside an EVP
# training step, the elementwise dropout and a bias-epilogue GEMM inside a MeLO training step (ViT-B, B = 32).
mkdir -p gpurun_out
for spec in "wgrad_kernel:evp:30" "hfreq_filter_kernel:evp:2" "dropout_kernel:melo:130" "gemm_bf16_sm100_kernel:melo:100"; do
  k=${spec%%:*}; rest=${spec#*:}; m=${rest%%:*}; s=${rest##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/full_r02s3_${m}_$k \
      python tools/profile_step.py --method $m --batch 32 --steps 1 > gpurun_out/ncu_full_s3_$k.log 2>&1
done
ls -la gpurun_out/full_r02s3_*.ncu-rep
