#!/bin/bash
# 2-GPU box: the 2-GPU bench line (parity block included), bounded
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=3000
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 30 > gpurun_out/r02p2b_bench_n2.json 2> gpurun_out/r02p2b_bench_n2.err
echo "bench rc=$?"; tail -c 3500 gpurun_out/r02p2b_bench_n2.json; tail -5 gpurun_out/r02p2b_bench_n2.err
