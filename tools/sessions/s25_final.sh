#!/bin/bash
# the whole GPU suite, no early stop
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q > gpurun_out/s25_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/s25_pytest.log
