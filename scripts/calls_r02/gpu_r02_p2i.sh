#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02p2i_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02p2i_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/sweep_p2p.py '{}' > gpurun_out/r02p2i_sweep.txt 2>&1
grep "us/update" gpurun_out/r02p2i_sweep.txt
