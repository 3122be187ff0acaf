#!/bin/bash
# time line of the data-parallel step at N = 2 (rank 0), global-batch and per-rank BN statistics
mkdir -p gpurun_out; export CENN_XR_TIMEOUT_S=20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/timeline.py > gpurun_out/s18_timeline_n2.txt 2> gpurun_out/s18_timeline_n2.err; echo "tl rc=$?"; head -3 gpurun_out/s18_timeline_n2.txt; tail -2 gpurun_out/s18_timeline_n2.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/timeline.py --bn-local > gpurun_out/s18_timeline_n2_local.txt 2> gpurun_out/s18_timeline_n2_local.err; echo "tl local rc=$?"; head -3 gpurun_out/s18_timeline_n2_local.txt; tail -2 gpurun_out/s18_timeline_n2_local.txt
timeout 300 python tools/timeline.py > gpurun_out/s18_timeline_n1.txt 2> gpurun_out/s18_timeline_n1.err; echo "tl n1 rc=$?"; head -3 gpurun_out/s18_timeline_n1.txt; tail -2 gpurun_out/s18_timeline_n1.txt
