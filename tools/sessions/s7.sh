#!/bin/bash
# GPU session 7: full tests, bench, per-op profile, ncu launch list + --set full captures of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log; tail -5 gpurun_out/s7_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err; echo "bench rc=$?"
timeout 200 python tools/profile_ops.py > gpurun_out/s7_ops.txt 2>&1; echo "ops rc=$?"
VARIANT=video B=64 timeout 200 python tools/profile_ops.py > gpurun_out/s7_ops_video.txt 2>&1; echo "ops video rc=$?"
timeout 200 python tools/timeline.py > gpurun_out/s7_timeline.txt 2>&1
timeout 120 python tools/ncu_step.py > gpurun_out/s7_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s7_launches.csv python tools/ncu_step.py > gpurun_out/s7_ncu_list.log 2>&1; echo "ncu list rc=$?"
for k in gather_gemm_kernel patch_dgrad_kernel wgrad_gemm_kernel; do
  timeout 120 python tools/ncu_step.py > gpurun_out/s7_ncu_plain.log 2>&1 &&
  STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 14 -o gpurun_out/s7_full_$k -f python tools/ncu_step.py > gpurun_out/s7_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 120 python tools/ncu_step.py > gpurun_out/s7_ncu_plain.log 2>&1 &&
STEPS=1 timeout 900 ncu --set full --clock-control none -k regex:'bn_|act_bwd|adam_bf16' -c 40 -o gpurun_out/s7_full_bw -f python tools/ncu_step.py > gpurun_out/s7_ncu_bw.log 2>&1; echo "ncu bw rc=$?"
ls -la gpurun_out/*.ncu-rep
