#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_sched.py 1024 '{"DQNB_BN_BIG": 128}' '{"DQNB_BN_BIG": 128, "DQNB_ST_DW": 3}' '{"DQNB_BN_BIG": 128, "DQNB_ST_DW": 4}' '{}' > gpurun_out/r02s_sweep.txt 2>&1
cat gpurun_out/r02s_sweep.txt
timeout 120 python scripts/trace_update.py 1024 '{"DQNB_BN_BIG": 128}' > gpurun_out/r02s_trace_bnbig.txt 2>&1
sed -n 26,40p gpurun_out/r02s_trace_bnbig.txt | cut -c1-150
