#!/bin/bash
# 2-GPU session: correctness of the real data-parallel data plane + a 2-GPU bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q -s > gpurun_out/s5_pytest_dp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest_dp.log; tail -12 gpurun_out/s5_pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s5_bench_n2.json 2> gpurun_out/s5_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/s5_bench_n2.err
