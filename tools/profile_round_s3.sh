mkdir -p gpurun_out
python tools/profile_step.py --method evp --batch 32 --by-site > gpurun_out/evp_step_by_site_b32.txt 2>&1
B="--steps 2 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline"
python bench.py $B > gpurun_out/bench_plain_r02g.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_r02g_bench_launches.csv python bench.py $B > gpurun_out/ncu_bench_run_g.log 2>&1
echo ncu_rc=$?
for k in wgrad_kernel hfreq_filter_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 14 -c 1 -o gpurun_out/full_r02_$k python tools/profile_step.py --method evp --batch 32 --steps 1 > gpurun_out/ncu_full_$k.log 2>&1
done
head -24 gpurun_out/evp_step_by_site_b32.txt | cut -c1-110
grep -c gvk gpurun_out/ncu_r02g_bench_launches.csv
ls -la gpurun_out/*.ncu-rep | tail -3
