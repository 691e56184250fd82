#!/bin/bash
# New-kernel validation pass: each group in its own process (a CUDA fault poisons its process only), then the bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
run() {  # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  local t0=$(date +%s)
  timeout $to python -m pytest "$@" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(( $(date +%s) - t0 ))s : $(tail -1 gpurun_out/$name.log)"
}
run n_fc_head 300 tests/test_gpu_b_kernels.py -k "fc_head"
run n_wgrad_batch 300 tests/test_gpu_b_kernels.py -k "wgrad_batch"
run n_mlp_chain 300 tests/test_gpu_b_kernels.py -k "mlp_chain"
run n_neighbours 300 tests/test_gpu_e_neighbours.py
run c_models_bf16 900 tests/test_gpu_c_models.py -k "bf16"
run c_models_f32 900 tests/test_gpu_c_models.py -k "f32 or share or relu or rejects"
run d_fullsize 900 tests/test_gpu_d_fullsize.py
for f in gpurun_out/n_*.log gpurun_out/c_*.log gpurun_out/d_*.log; do echo "---- $f"; grep -E "^(FAILED|ERROR)|Error|error|assert " $f | head -12; done
if [ "$1" == "bench" ]; then
  timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; tail -c 5000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
