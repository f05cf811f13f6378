#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/sweep_p2p.py '{"DQNB_P2P_FENCE_ALL": 1}' '{}' > gpurun_out/r02p2h_sweep.txt 2>&1
grep "us/update" gpurun_out/r02p2h_sweep.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 scripts/trace_update.py 1024 > gpurun_out/r02p2h_trace.txt 2>&1
grep -E "REDUCE|P2P|ADAM|graph replay" gpurun_out/r02p2h_trace.txt | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02p2h_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02p2h_tests.log
