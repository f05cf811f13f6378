#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02r_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02r_tests.log
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_ADAM_SPLIT": 0}' '{}' '{"DQNB_ADAM_SPLIT": 0}' > gpurun_out/r02r_sweep.txt 2>&1
cat gpurun_out/r02r_sweep.txt
timeout 120 python scripts/trace_update.py 1024 > gpurun_out/r02r_trace.txt 2>&1
grep -E "REDUCE|ADAM|graph replay" gpurun_out/r02r_trace.txt | cut -c1-150
