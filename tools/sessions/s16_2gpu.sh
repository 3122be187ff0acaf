#!/bin/bash
# where does the data-parallel overhead come from (N = 2): NCCL CTA budget, bucket format, early Adam; bn_local parity re-run
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-video-block --no-local-bn-block --no-cpu-baseline > gpurun_out/s16_n2_$tag.json 2> gpurun_out/s16_n2_$tag.err; echo "n2 $tag rc=$? $(python -c "import json;d=json.loads(open('gpurun_out/s16_n2_$tag.json').read().strip().splitlines()[-1]);print(d['ms_per_step'], d['value'])")"; }
run default X=1
run ctas4 NCCL_MAX_CTAS=4
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run fp32b CENN_FP32_BUCKETS=1
run noearly CENN_NO_EARLY_ADAM=1
run default2 X=1
timeout 900 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q -k "bn-local or env6 or env7" > gpurun_out/s16_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -5 gpurun_out/s16_pytest_dp.log
