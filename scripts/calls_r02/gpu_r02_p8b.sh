#!/bin/bash
# 8-GPU box: the 8-rank tests (cfg2 x8 + cfg5 as specified; cfg4: 64 workers over 8 ranks)
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=10000
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "eight" > gpurun_out/r02p8b_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r02p8b_tests.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 scripts/cfg4_rollout.py > gpurun_out/r02p8b_cfg4.txt 2>&1
grep -E "CFG4|ok" gpurun_out/r02p8b_cfg4.txt | cut -c1-1200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 scripts/multi_gpu_check.py cfg5 > gpurun_out/r02p8b_cfg5.txt 2>&1
grep -E "cfg5|MULTI" gpurun_out/r02p8b_cfg5.txt | cut -c1-1500
