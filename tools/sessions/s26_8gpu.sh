#!/bin/bash
# 8 GPUs, short: the image line with the peer-memory gradient path complete (leftover ranges + losses)
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=30
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-local-bn-block > gpurun_out/s26_bench_n8.json 2> gpurun_out/s26_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s26_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'video', (d.get('video') or {}).get('value'), (d.get('video') or {}).get('ms_per_step'))
print(d['config']['losses_last_step'])
PY
