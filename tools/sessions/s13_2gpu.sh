#!/bin/bash
# 2 GPUs: DP parity incl. optional branches and pull-mode exchange; bench after the all-reduce ordering change
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 1200 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s13_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -15 gpurun_out/s13_pytest_dp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s13_bench_n2.json 2> gpurun_out/s13_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/s13_bench_n2.json | head -c 600
timeout 600 python -m pytest tests/test_fused_branches_gpu.py -m gpu -q -x > gpurun_out/s13_branches.log 2>&1; echo "branches rc=$?"; tail -15 gpurun_out/s13_branches.log
