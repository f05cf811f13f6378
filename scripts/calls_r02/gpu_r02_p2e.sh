#!/bin/bash
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=3000
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/sweep_p2p.py '{"DQNB_P2P_PUSH": 0}' '{}' > gpurun_out/r02p2e_sweep.txt 2>&1
cat gpurun_out/r02p2e_sweep.txt | grep -v "^\*\|OMP_NUM"
