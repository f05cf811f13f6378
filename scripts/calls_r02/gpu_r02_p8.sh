#!/bin/bash
# 8-GPU box: data-parallel parity at the BASELINE shapes (cfg2 x8, cfg5 as specified), cfg4 with 8 ranks, the 8-GPU bench line
mkdir -p gpurun_out
export DQNB_P2P_TIMEOUT_MS=10000
nvidia-smi -L | wc -l
timeout 500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "eight" > gpurun_out/r02p8_tests.log 2>&1
echo "tests rc=$?"; tail -6 gpurun_out/r02p8_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 300 --warmup 30 > gpurun_out/r02p8_bench_n8.json 2> gpurun_out/r02p8_bench_n8.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02p8_bench_n8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "windows_ms_per_step")}, d["e2e"]["value"], d["parity"]["ok"], d["wide_mlp"])
PY
tail -3 gpurun_out/r02p8_bench_n8.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29553 scripts/trace_update.py 1024 > gpurun_out/r02p8_trace.txt 2>&1
grep -E "REDUCE|P2P|ADAM|graph replay" gpurun_out/r02p8_trace.txt | cut -c1-260
