#!/bin/bash
# GPU session 3 (round 2): full GPU tests on the cleaned-up build, bench with / without PDL, op-level path, time line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log; tail -4 gpurun_out/s3_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
CENN_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video-block > gpurun_out/s3_bench_nopdl.json 2> gpurun_out/s3_bench_nopdl.err; echo "bench nopdl rc=$?"
timeout 300 python bench.py --path oplevel --steps 5 --warmup 3 > gpurun_out/s3_bench_oplevel.json 2> gpurun_out/s3_bench_oplevel.err; echo "oplevel rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err; echo "ref rc=$?"
timeout 200 python tools/timeline.py > gpurun_out/s3_timeline.txt 2>&1; echo "timeline rc=$?"
timeout 200 python tools/profile_ops.py > gpurun_out/s3_ops.txt 2>&1; echo "ops rc=$?"
