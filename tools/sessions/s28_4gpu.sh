#!/bin/bash
# 4 GPUs: the world size the round-end scaling run uses and no session of this round had exercised
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=30
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-local-bn-block > gpurun_out/s28_bench_n4.json 2> gpurun_out/s28_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s28_bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'video', (d.get('video') or {}).get('value'), (d.get('video') or {}).get('ms_per_step'))
print(d['config']['losses_last_step'])
PY
