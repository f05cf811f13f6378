#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu > gpurun_out/r02u_tests_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -4 gpurun_out/r02u_tests_gemm.log
timeout 120 python scripts/perf_gemm_ts.py > gpurun_out/r02u_probe.txt 2>&1; cat gpurun_out/r02u_probe.txt
