#!/bin/bash
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=15
run() { tag=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block > gpurun_out/s6_bench_n2_$tag.json 2> gpurun_out/s6_bench_n2_$tag.err; echo "bench n2 $tag rc=$?"; }
run default X=1
run nopdl CENN_PDL=0
run fold CENN_DP_BN_FOLD=1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s6_bench_n2_full.json 2> gpurun_out/s6_bench_n2_full.err; echo "bench n2 full rc=$?"
timeout 600 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s6_pytest_dp.log 2>&1; tail -3 gpurun_out/s6_pytest_dp.log
for f in gpurun_out/s6_bench_n2_*.json; do echo $f; head -c 300 $f; echo; done
