#!/bin/bash
# noiseGen / conditionAdv in the fused executor: new parity tests + regression of the main fused tests + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fused_branches_gpu.py -m gpu -q -x > gpurun_out/s12_branches.log 2>&1; echo "branches rc=$?"; tail -30 gpurun_out/s12_branches.log
timeout 900 python -m pytest tests/test_fused_gpu.py -m gpu -q -x -k "matches_oracle or discriminator_blocks" > gpurun_out/s12_fused.log 2>&1; echo "fused rc=$?"; tail -5 gpurun_out/s12_fused.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/s12_bench.json
