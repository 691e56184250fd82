#!/bin/bash
# compute-sanitizer over the smoke shapes and the fused-kernel tests: memcheck (out-of-bounds / misaligned accesses),
# racecheck (shared-memory hazards), synccheck (barrier misuse).  Summaries -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 900 $CS --tool $tool --error-exitcode 7 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$? : $(grep -E 'ERROR SUMMARY|smoke ok' gpurun_out/sanitize_${tool}_smoke.log | tr '\n' ' ')"
done
timeout 1500 $CS --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_f_fused.py tests/test_gpu_b_kernels.py -m gpu -q -x -p no:cacheprovider \
  -k "fused_forward_layer_and_pool or adjoint_layer or head_du or fused_adam or pool or scores" > gpurun_out/sanitize_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_memcheck_kernels.log | tr '\n' ' ')"
timeout 1500 $CS --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_f_fused.py -m gpu -q -x -p no:cacheprovider \
  -k "head_du or (fused_forward_layer_and_pool and shape1) or (adjoint_layer and shape1)" > gpurun_out/sanitize_racecheck_fused.log 2>&1
echo "racecheck fused rc=$? : $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY|passed|failed' gpurun_out/sanitize_racecheck_fused.log | tr '\n' ' ')"
timeout 900 $CS --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitize_racecheck_smoke.log 2>&1
echo "racecheck smoke rc=$? : $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY|smoke ok' gpurun_out/sanitize_racecheck_smoke.log | tr '\n' ' ')"
