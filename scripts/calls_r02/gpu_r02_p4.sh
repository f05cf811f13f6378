#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02p4_bench_n4.json 2> gpurun_out/r02p4_bench_n4.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p4_bench_n4.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "windows_ms_per_step")}, d["e2e"]["value"], d["parity"]["ok"], d["wide_mlp"]["value"])
PY
