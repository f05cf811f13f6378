#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_BN_BIG": 128}' '{"DQNB_BN_BIG": 128, "DQNB_DW_BIG_MAX_CTAS": 80}' '{"DQNB_DW_BIG_MAX_CTAS": 80}' '{}' > gpurun_out/r02t_sweep.txt 2>&1
cat gpurun_out/r02t_sweep.txt
timeout 120 python scripts/trace_update.py 1024 '{"DQNB_BN_BIG": 128, "DQNB_DW_BIG_MAX_CTAS": 80}' > gpurun_out/r02t_trace.txt 2>&1
sed -n 26,38p gpurun_out/r02t_trace.txt | cut -c1-150
