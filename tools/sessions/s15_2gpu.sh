#!/bin/bash
# per-rank BN statistics (cfg.bn_local): DP parity + 2-GPU bench with the local_bn sub-block
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 1500 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s15_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -25 gpurun_out/s15_pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s15_bench_n2.json 2> gpurun_out/s15_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/s15_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s15_bench_n2.json').read().strip().splitlines()[-1])
print('sync', d['value'], d['ms_per_step'], 'video', d['video']['value'], d['video']['ms_per_step'])
print('local', d.get('local_bn'))
PY
