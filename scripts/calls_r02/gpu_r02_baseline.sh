#!/bin/bash
# round 2, first call: baseline of the round-1 build on today's box + measurements the round-1 verdict asked for.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
python bench.py --steps 500 --warmup 50 --cpu-seconds 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err
tail -c 600 gpurun_out/r02a_bench_n1.json
# dense TF32 / bf16 peaks through torch (cuBLAS), 8192^3
python - <<'PY' > gpurun_out/r02a_tf32_peak.json
import json, torch
torch.backends.cuda.matmul.allow_tf32 = True
out = {}
for name, dt in (("tf32", torch.float32), ("bf16", torch.bfloat16)):
    a = torch.randn(8192, 8192, device="cuda", dtype=dt); b = torch.randn(8192, 8192, device="cuda", dtype=dt)
    for _ in range(5): a @ b
    best = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0.record()
    for _ in range(200): c = a @ b
    e1.record(); torch.cuda.synchronize()
    out[name] = {"burst_tflops": 2 * 8192**3 / best / 1e9, "sustained_tflops": 2 * 8192**3 * 200 / e0.elapsed_time(e1) / 1e9}
print(json.dumps(out))
PY
cat gpurun_out/r02a_tf32_peak.json
# ncu --set full of the HBM/L2-bound kernels of the update (graph replay -> kernel nodes are profiled one by one)
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'adam_kernel|reduce_kernel|gather_kernel|colsum_kernel|head_bwd_w_kernel|critic_head_kernel|actor_head_bwd_kernel|head_fwd_kernel' \
  -s 60 -c 24 -o gpurun_out/r02a_small_kernels -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu_small.log 2>&1
tail -3 gpurun_out/r02a_ncu_small.log
python scripts/trace_update.py 1024 > gpurun_out/r02a_trace_b1024.txt 2>&1
tail -5 gpurun_out/r02a_trace_b1024.txt
ls -la gpurun_out | head -40
