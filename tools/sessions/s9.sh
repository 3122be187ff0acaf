#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log; grep -E "^E  |passed|failed" gpurun_out/s9_pytest.log | head -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; echo "bench rc=$?"
CENN_THIN_ONE_CTA=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-video-block > gpurun_out/s9_bench_thin1.json 2> gpurun_out/s9_bench_thin1.err; echo "bench thin1 rc=$?"
timeout 200 python tools/profile_ops.py > gpurun_out/s9_ops.txt 2>&1; cp gpurun_out/ops_image_b256.txt gpurun_out/s9_ops_image.txt
for f in gpurun_out/s9_bench.json gpurun_out/s9_bench_thin1.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
