#!/bin/bash
# 8 GPUs: parity of the sharded path at world 8, final bench line (video + local_bn sub-blocks), NCCL-bucket A/B, time line
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=30
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 tools/dp_parity.py --variant image --nB 4000 --per-rank 2 --steps 3 > gpurun_out/s22_parity_n8.log 2> gpurun_out/s22_parity_n8.err; echo "parity n8 rc=$?"; grep DP_PARITY gpurun_out/s22_parity_n8.log | cut -c1-900
timeout 400 $T --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s22_bench_n8.json 2> gpurun_out/s22_bench_n8.err; echo "bench n8 rc=$?"
timeout 200 env CENN_NO_SHARD_ADAM=1 $T --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-video-block --no-local-bn-block > gpurun_out/s22_bench_n8_nccl.json 2> gpurun_out/s22_bench_n8_nccl.err; echo "bench n8 nccl rc=$?"
timeout 200 $T --master-port 29514 tools/timeline.py > gpurun_out/s22_timeline_n8.txt 2> gpurun_out/s22_timeline_n8.err; echo "tl rc=$?"
python - <<'PY'
import json
for f in ('s22_bench_n8','s22_bench_n8_nccl'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'video', (d.get('video') or {}).get('value'), 'local', (d.get('local_bn') or {}).get('value'), ((d.get('local_bn') or {}).get('video') or {}).get('value'))
        print('   losses', d['config']['losses_last_step'])
    except Exception as e: print(f, 'ERR', e)
PY
head -2 gpurun_out/s22_timeline_n8.txt
