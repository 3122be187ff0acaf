#!/bin/bash
# BN-backward sums from the dgrad epilogues: parity + A/B bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_fused_branches_gpu.py -m gpu -q -x > gpurun_out/s14_fused.log 2>&1; echo "fused rc=$?"; tail -25 gpurun_out/s14_fused.log
for tag in epi noepi epi2 noepi2; do
  if [ "${tag:0:2}" = "no" ]; then export CENN_NO_BWD_EPI=1; else unset CENN_NO_BWD_EPI; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video-block > gpurun_out/s14_bench_$tag.json 2> gpurun_out/s14_bench_$tag.err; echo "bench $tag rc=$?"; head -c 230 gpurun_out/s14_bench_$tag.json; echo
done
unset CENN_NO_BWD_EPI
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload video > gpurun_out/s14_bench_video.json 2> gpurun_out/s14_bench_video.err; echo "video rc=$?"; head -c 230 gpurun_out/s14_bench_video.json; echo
CENN_NO_BWD_EPI=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload video > gpurun_out/s14_bench_video_noepi.json 2> gpurun_out/s14_bench_video_noepi.err; echo "video noepi rc=$?"; head -c 230 gpurun_out/s14_bench_video_noepi.json; echo
