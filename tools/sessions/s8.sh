#!/bin/bash
# GPU session 8: full tests, bench, per-op profiles, ncu launch list + --set full captures exported as CSV (the .ncu-rep files stay on the box)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log; grep -E "^E  |passed|failed" gpurun_out/s8_pytest.log | head -12
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "bench rc=$?"
CENN_BN_BWD_3LAUNCH=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video-block > gpurun_out/s8_bench_bn3.json 2> gpurun_out/s8_bench_bn3.err; echo "bench bn3 rc=$?"
timeout 200 python tools/profile_ops.py > gpurun_out/s8_ops.txt 2>&1; echo "ops rc=$?"
VARIANT=video B=64 timeout 200 python tools/profile_ops.py > gpurun_out/s8_ops_video.txt 2>&1; echo "ops video rc=$?"
timeout 200 python tools/timeline.py > gpurun_out/s8_timeline.txt 2>&1
R=/tmp/ncu_reps; mkdir -p $R
timeout 120 python tools/ncu_step.py > gpurun_out/s8_ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s8_launches.csv python tools/ncu_step.py > gpurun_out/s8_ncu_list.log 2>&1; echo "ncu list rc=$?"
cap() { k=$1; n=$2; tag=$3
  timeout 120 python tools/ncu_step.py > gpurun_out/s8_ncu_plain.log 2>&1 &&
  STEPS=1 timeout 900 ncu --set full --clock-control none -k regex:"$k" -c $n -o $R/$tag -f python tools/ncu_step.py > gpurun_out/s8_ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
  ncu -i $R/$tag.ncu-rep --page raw --csv > gpurun_out/s8_full_$tag.csv 2>/dev/null; ls -la gpurun_out/s8_full_$tag.csv; }
cap gather_gemm_kernel 12 gather
cap patch_dgrad_kernel 8 patch
cap wgrad_gemm_kernel 8 wgrad
cap 'bn_|act_bwd|adam_bf16' 30 bw
du -sh gpurun_out
