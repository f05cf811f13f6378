#!/bin/bash
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=3000
DQNB_P2P_PUSH=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 scripts/trace_update.py 1024 > gpurun_out/r02p2f_trace_pull.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 scripts/trace_update.py 1024 > gpurun_out/r02p2f_trace_push.txt 2>&1
grep -E "REDUCE|P2P|ADAM|graph replay" gpurun_out/r02p2f_trace_pull.txt | cut -c1-260
grep -E "REDUCE|P2P|ADAM|graph replay" gpurun_out/r02p2f_trace_push.txt | cut -c1-260
