#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -m gpu -q -k "100_steps or clip_mode or frame_mode or byte_image" > gpurun_out/s27_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s27_pytest.log
