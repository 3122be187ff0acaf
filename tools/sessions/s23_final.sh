#!/bin/bash
# final 1-GPU verification of the round: smoke, the whole GPU suite, the default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s23_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s23_smoke.log
timeout 400 python bench.py > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/s23_bench.json; echo
timeout 1500 python -m pytest tests/ -m gpu -q -x > gpurun_out/s23_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/s23_pytest.log
