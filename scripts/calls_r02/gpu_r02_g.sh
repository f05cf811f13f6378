#!/bin/bash
# full GPU suite + the default bench line of the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02g_tests.log
timeout 900 python bench.py --steps 300 --warmup 30 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02g_bench_n1.json; tail -5 gpurun_out/r02g_bench_n1.err
