#!/bin/bash
# 2-GPU box: data-parallel parity (small + cfg3 as specified) and the 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02p2_smi.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02p2_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r02p2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 30 > gpurun_out/r02p2_bench_n2.json 2> gpurun_out/r02p2_bench_n2.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02p2_bench_n2.json; tail -5 gpurun_out/r02p2_bench_n2.err
