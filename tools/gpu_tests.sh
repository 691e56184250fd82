#!/bin/bash
# Runs the GPU parity tests in separate processes (a CUDA fault poisons its process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  local t0=$(date +%s)
  timeout $to python -m pytest "$@" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(( $(date +%s) - t0 ))s : $(tail -1 gpurun_out/$name.log)"
}
run a_integer 600 tests/test_gpu_a_integer.py
run b_kernels_f32 900 tests/test_gpu_b_kernels.py -k "f32"
run b_kernels_bf16 900 tests/test_gpu_b_kernels.py -k "bf16"
run b_kernels_rest 900 tests/test_gpu_b_kernels.py -k "not f32 and not bf16"
run c_models 900 tests/test_gpu_c_models.py
run d_fullsize 900 tests/test_gpu_d_fullsize.py
run e_neighbours 600 tests/test_gpu_e_neighbours.py
run f_fused 600 tests/test_gpu_f_fused.py
for f in gpurun_out/[a-f]_*.log; do echo "---- $f"; grep -E "^(FAILED|ERROR)|Error|error|assert " $f | head -12; done
