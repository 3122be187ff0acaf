#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log; tail -15 gpurun_out/s4_pytest.log
timeout 300 python bench.py --workload deeper --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/s4_bench_deeper.json 2> gpurun_out/s4_bench_deeper.err; echo "deeper rc=$?"; tail -2 gpurun_out/s4_bench_deeper.err
