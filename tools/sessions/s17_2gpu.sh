#!/bin/bash
# N = 2: NCCL CTA floor and bucket granularity
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-local-bn-block --no-cpu-baseline > gpurun_out/s17_n2_$tag.json 2> gpurun_out/s17_n2_$tag.err; echo "n2 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s17_n2_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])")"; }
run default X=1
run min32 NCCL_MIN_CTAS=32
run min64 NCCL_MIN_CTAS=64
run g18 CENN_G_BUCKET_LOG2=18
run g16 CENN_G_BUCKET_LOG2=16
run g18d16 CENN_G_BUCKET_LOG2=18 CENN_D_BUCKET_LOG2=16
run g18min32 CENN_G_BUCKET_LOG2=18 NCCL_MIN_CTAS=32
run default2 X=1
grep -h "NCCL INFO.*hannel\|nChannels" gpurun_out/s17_n2_default.err | head -5
