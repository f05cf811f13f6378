#!/bin/bash
# ncu passes of the profiling recipe (B200_PROFILING.md): launch list with per-launch device time,
# then one --set full capture of the dominant kernel.
mkdir -p gpurun_out
# 62 kernels per update (+ a few setup kernels): skip the first two updates, list the third
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 62 --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_update.py 3 0 > gpurun_out/profile_run.log 2>&1
tail -3 gpurun_out/profile_run.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 6 -o gpurun_out/prof_gemm \
    python scripts/profile_update.py 2 0 > gpurun_out/profile_run2.log 2>&1
tail -3 gpurun_out/profile_run2.log
ls -la gpurun_out/
