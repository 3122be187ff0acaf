#!/bin/bash
# sharded reduce + Adam over peer memory for the big generator blocks: parity, A/B bench, time line (N = 2)
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 900 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q -x -k "env10 or env11 or env0 or env1" > gpurun_out/s20_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -12 gpurun_out/s20_pytest_dp.log
r2() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-local-bn-block --no-cpu-baseline > gpurun_out/s20_n2_$tag.json 2> gpurun_out/s20_n2_$tag.err; echo "n2 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s20_n2_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['e2e']['value'])")"; tail -2 gpurun_out/s20_n2_$tag.err | cut -c1-300; }
r2 shard X=1
r2 nccl CENN_NO_SHARD_ADAM=1
r2 shard2 X=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/timeline.py > gpurun_out/s20_timeline_n2.txt 2> gpurun_out/s20_timeline_n2.err; echo "tl rc=$?"; head -2 gpurun_out/s20_timeline_n2.txt; awk '$3==4 || $2 ~ /join_adam|gradG_sync|finish/' gpurun_out/s20_timeline_n2.txt | tail -12
