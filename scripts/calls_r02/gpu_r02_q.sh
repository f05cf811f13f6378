#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/r02q_tests_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -5 gpurun_out/r02q_tests_gemm.log
timeout 900 python -m pytest tests/test_gpu_update.py tests/test_gpu_configs.py tests/test_golden.py tests/test_gpu_share.py -x -q -m gpu > gpurun_out/r02q_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02q_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_TS_MIN_KB": 9999}' '{"DQNB_TS_MIN_KB": 4}' '{"DQNB_TS_MIN_KB": 16}' '{}' > gpurun_out/r02q_sweep.txt 2>&1
cat gpurun_out/r02q_sweep.txt
timeout 120 python scripts/trace_update.py 1024 > gpurun_out/r02q_trace.txt 2>&1
timeout 120 python scripts/perf_gemm_ts.py > gpurun_out/r02q_probe.txt 2>&1; cat gpurun_out/r02q_probe.txt
