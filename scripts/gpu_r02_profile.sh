#!/bin/bash
# round-2 profiling passes (B200_PROFILING.md): launch list of the bench command, --set full captures of the GEMM and of
# the memory-bound kernels, the in-kernel timeline, the default bench line.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02z_smi.txt
timeout 600 python bench.py --steps 500 --warmup 50 > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02z_bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err
echo "reference rc=$?"; tail -c 300 gpurun_out/r02z_bench_reference.json
timeout 120 python scripts/trace_update.py 1024 > gpurun_out/r02z_trace.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv \
    python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > gpurun_out/r02z_ncu_launch_run.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02z_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 120 -c 38 -o gpurun_out/r02z_gemm -f \
    python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > gpurun_out/r02z_ncu_gemm_run.log 2>&1
echo "gemm capture rc=$?"
# raw metrics as CSV on the box (the .ncu-rep of 38 launches is ~48 MB; gpurun brings back at most 64 MiB)
ncu -i gpurun_out/r02z_gemm.ncu-rep --page raw --csv > gpurun_out/r02z_gemm_raw.csv 2>/dev/null; rm -f gpurun_out/r02z_gemm.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'adam_kernel|reduce_kernel|gather_kernel|colsum_kernel|head_bwd_w_kernel|critic_head_kernel|actor_head_bwd_kernel|head_fwd_kernel' \
    -s 40 -c 20 -o gpurun_out/r02z_small -f python bench.py --steps 2 --warmup 3 --windows 1 --no-extras --no-cpu-baseline > gpurun_out/r02z_ncu_small_run.log 2>&1
echo "small capture rc=$?"
ncu -i gpurun_out/r02z_small.ncu-rep --page raw --csv > gpurun_out/r02z_small_raw.csv 2>/dev/null; rm -f gpurun_out/r02z_small.ncu-rep
ls -la gpurun_out/r02z_*
