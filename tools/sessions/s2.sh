#!/bin/bash
# GPU session 2 (round 2): tests + bench variants + CTA-pair probe.  Run from the repo root under gpurun.
mkdir -p gpurun_out; export CENN_THIN_IM2COL=${CENN_THIN_IM2COL-1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log; tail -4 gpurun_out/s2_pytest.log
for v in default thinold frac85 bn3; do
  case $v in default) E="X=1";; thinold) E="CENN_THIN_IM2COL=1";; frac85) E="CENN_BN_FRAC=0.85";; bn3) E="CENN_BN_BWD_3LAUNCH=1";; esac
  env $E timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video-block > gpurun_out/s2_bench_$v.json 2> gpurun_out/s2_bench_$v.err; echo "bench $v rc=$?"
done
timeout 300 python bench.py --workload video --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s2_bench_video.json 2> gpurun_out/s2_bench_video.err; echo "video rc=$?"
(timeout 60 python tools/gemm2sm_probe.py 256 256 64; timeout 60 python tools/gemm2sm_probe.py 256 256 512; timeout 90 python tools/gemm2sm_probe.py 4096 4096 4096 20; timeout 90 python tools/gemm2sm_probe.py 8192 8192 4096 20; timeout 90 python tools/gemm_big_probe.py) > gpurun_out/s2_probe.log 2>&1
cat gpurun_out/s2_probe.log
