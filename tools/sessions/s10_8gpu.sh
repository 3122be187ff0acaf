#!/bin/bash
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s10_bench_n8.json 2> gpurun_out/s10_bench_n8.err; echo "bench n8 rc=$?"; tail -2 gpurun_out/s10_bench_n8.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload deeper --steps 10 --warmup 4 > gpurun_out/s10_bench_deeper_n8.json 2> gpurun_out/s10_bench_deeper_n8.err; echo "deeper n8 rc=$?"; tail -2 gpurun_out/s10_bench_deeper_n8.err | cut -c1-300
for f in gpurun_out/s10_bench_n8.json gpurun_out/s10_bench_deeper_n8.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('video',{}).get('value'), d.get('e2e_bytes',{}).get('value'))"; done
