#!/bin/bash
# compute-sanitizer passes over the kernel unit tests and one small update (the reference has no
# sanitizer coverage; its benign-by-neglect races are on the host side).  Run under gpurun, ~3-5 GPU-minutes:
#   gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh'
# memcheck: out-of-bounds / misaligned global+shared accesses (TMA transfers are not visible to it);
# racecheck: shared-memory hazards between the warp roles (mbarrier-ordered smem reuse in the GEMM epilogue);
# synccheck: divergent / invalid barrier use (named barrier 1, cluster barriers).
mkdir -p gpurun_out
SMALL='import __graft_entry__ as g; g.smoke()'
for tool in memcheck racecheck synccheck; do
  echo "== $tool: smoke()"
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 7 python -c "$SMALL" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_smoke.log) summaries"; tail -2 gpurun_out/sanitize_${tool}_smoke.log
done
echo "== memcheck: gemm unit tests (tile / ring variants)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu -k "tile_and_ring or exact or tensor_memory" \
  > gpurun_out/sanitize_memcheck_gemm.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/sanitize_memcheck_gemm.log
