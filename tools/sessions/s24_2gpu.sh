#!/bin/bash
# peer all-reduce of the leftover gradient ranges + mailbox loss exchange: parity subset and bench at N = 2
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 900 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q -k "env0 or env1 or env5 or env8" > gpurun_out/s24_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -6 gpurun_out/s24_pytest_dp.log
r2() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-local-bn-block --no-cpu-baseline > gpurun_out/s24_n2_$tag.json 2> gpurun_out/s24_n2_$tag.err; echo "n2 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s24_n2_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'], d['e2e']['value'])")"; }
r2 peer X=1
r2 nopeer CENN_NO_PEER_AR=1
