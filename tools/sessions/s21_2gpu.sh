#!/bin/bash
# the whole data-parallel parity suite at 2 GPUs (incl. the sharded reduce + Adam cases at nBottleneck 4000)
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 1500 python -m pytest tests/test_dp_multi_gpu.py -m gpu -q > gpurun_out/s21_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -15 gpurun_out/s21_pytest_dp.log
