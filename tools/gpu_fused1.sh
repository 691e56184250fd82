#!/bin/bash
# fused layer kernel: its parity tests, then the micro-benchmark sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_fused.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/fused_tests.log 2>&1
echo "== tests exit=$? : $(tail -1 gpurun_out/fused_tests.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/fused_tests.log | head -20
timeout 600 python tools/bench_fused.py > gpurun_out/bench_fused.log 2>&1
echo "bench rc=$?"; cat gpurun_out/bench_fused.log | tail -40
