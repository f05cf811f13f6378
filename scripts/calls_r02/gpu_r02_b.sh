#!/bin/bash
# round 2, call b: parity of the fused epilogue column sums + fused target/loss head, A/B timing, reference-main binary
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_update.py tests/test_gpu_gemm.py tests/test_golden.py -x -q -m gpu > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02b_tests.log
timeout 300 python scripts/sweep_sched.py 1024 '{"DQNB_FUSE_COLSUM": 0, "DQNB_FUSE_TL": 0}' '{"DQNB_FUSE_COLSUM": 1, "DQNB_FUSE_TL": 0}' '{"DQNB_FUSE_COLSUM": 0, "DQNB_FUSE_TL": 1}' '{}' > gpurun_out/r02b_sweep.txt 2>&1
cat gpurun_out/r02b_sweep.txt
timeout 300 python -m pytest tests/test_host.py -x -q -m gpu > gpurun_out/r02b_host.log 2>&1
echo "host rc=$?"; tail -15 gpurun_out/r02b_host.log
python scripts/trace_update.py 1024 > gpurun_out/r02b_trace_b1024.txt 2>&1
