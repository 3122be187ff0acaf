#!/bin/bash
# chunked buckets + comm-stream priority + deferred dead dgrad: parity and A/B (N = 1 and N = 2)
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
r1() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-video-block --no-cpu-baseline > gpurun_out/s19_n1_$tag.json 2> gpurun_out/s19_n1_$tag.err; echo "n1 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s19_n1_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])")"; }
r2() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-local-bn-block --no-cpu-baseline > gpurun_out/s19_n2_$tag.json 2> gpurun_out/s19_n2_$tag.err; echo "n2 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s19_n2_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])")"; }
r1 default X=1
r1 inline CENN_DEAD_DGRAD_INLINE=1
r1 default2 X=1
r1 inline2 CENN_DEAD_DGRAD_INLINE=1
r2 default X=1
r2 inline CENN_DEAD_DGRAD_INLINE=1
r2 chunks1 CENN_BUCKET_CHUNKS=1
r2 chunks8 CENN_BUCKET_CHUNKS=8
r2 prio0 CENN_COMM_PRIO=0
r2 old CENN_COMM_PRIO=0 CENN_BUCKET_CHUNKS=1 CENN_DEAD_DGRAD_INLINE=1
r2 default2 X=1
timeout 1500 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s19_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -5 gpurun_out/s19_pytest_dp.log
timeout 900 python -m pytest tests/test_fused_gpu.py -m gpu -q -x -k "matches_oracle or byte_image or benchmark_shapes" > gpurun_out/s19_fused.log 2>&1; echo "fused rc=$?"; tail -3 gpurun_out/s19_fused.log
