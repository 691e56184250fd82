#!/bin/bash
# Runs the GPU parity tests in separate processes (a CUDA fault poisons its process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  timeout $to python -m pytest "$@" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? : $(tail -1 gpurun_out/$name.log)"
}
run a_integer 600 tests/test_gpu_a_integer.py
run b_aggregate 600 tests/test_gpu_b_kernels.py -k aggregate
run b_linear_f32 600 tests/test_gpu_b_kernels.py -k "test_linear and f32"
run b_wgrad_f32 600 tests/test_gpu_b_kernels.py -k "test_wgrad and f32"
run b_pool_scores 600 tests/test_gpu_b_kernels.py -k "pool or scores"
run c_models_f32 900 tests/test_gpu_c_models.py -k "f32 or share or relu or rejects"
run b_linear_bf16 600 tests/test_gpu_b_kernels.py -k "test_linear and bf16"
run b_wgrad_bf16 600 tests/test_gpu_b_kernels.py -k "test_wgrad and bf16"
run b_tc_big 600 tests/test_gpu_b_kernels.py -k "tensor_core"
run c_models_bf16 900 tests/test_gpu_c_models.py -k "bf16"
for f in gpurun_out/*.log; do echo "---- $f"; grep -E "^(FAILED|ERROR)|Error|error|assert " $f | head -12; done
