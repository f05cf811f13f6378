#!/bin/bash
# 8-GPU box, final build: cfg2 x8 + cfg5-as-specified parity test, the 8-GPU bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "eight_rank_wide" > gpurun_out/r02p8c_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r02p8c_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02p8c_bench_n8.json 2> gpurun_out/r02p8c_bench_n8.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p8c_bench_n8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "windows_ms_per_step")}, d["e2e"]["value"], d["parity"]["ok"], d["wide_mlp"]["value"], d["wide_mlp"]["ms_per_step"], d["act_path"])
PY
tail -2 gpurun_out/r02p8c_bench_n8.err
