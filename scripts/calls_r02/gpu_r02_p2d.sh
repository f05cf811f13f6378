#!/bin/bash
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=3000
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 300 --warmup 30 > gpurun_out/r02p2d_bench_n2.json 2> gpurun_out/r02p2d_bench_n2.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p2d_bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "windows_ms_per_step")}, d["e2e"]["value"], d["parity"]["ok"], d.get("act_path"), d["wide_mlp"]["ms_per_step"])
PY
tail -3 gpurun_out/r02p2d_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02p2d_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02p2d_tests.log
