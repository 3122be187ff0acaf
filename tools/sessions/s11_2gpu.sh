#!/bin/bash
# push-mode exchange A/B at 2 GPUs + DP parity tests
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=15
run() { tag=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-cpu-baseline > gpurun_out/s11_bench_n2_$tag.json 2> gpurun_out/s11_bench_n2_$tag.err; echo "bench n2 $tag rc=$?"; }
run push X=1
run pull CENN_XR_PULL=1
run push2 X=1
run pull2 CENN_XR_PULL=1
timeout 600 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s11_pytest_dp.log 2>&1; tail -3 gpurun_out/s11_pytest_dp.log
for f in gpurun_out/s11_bench_n2_*.json; do echo $f; head -c 200 $f; echo; done
