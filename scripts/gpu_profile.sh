#!/bin/bash
# ncu passes of the profiling recipe (B200_PROFILING.md): launch list with per-launch device time
# (every launch of a 3-update run; the last update = the last 53 rows), then one --set full capture of
# the dominant kernel.  Eager mode (use_graph=0) so that every kernel is a separate, attributable launch.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_update.py 3 0 > gpurun_out/profile_run.log 2>&1
tail -2 gpurun_out/profile_run.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 10 -o gpurun_out/prof_gemm \
    python scripts/profile_update.py 2 0 > gpurun_out/profile_run2.log 2>&1
tail -2 gpurun_out/profile_run2.log
