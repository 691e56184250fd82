#!/bin/bash
# ncu --set full of the fp32-parity (split operand) kernels of tools/bench_split.py
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"linear_ws_kernel|wgrad_tall_kernel|split_f16_kernel|amax_kernel|split_reduce_bias_scaled" -c 9 -f -o gpurun_out/prof_split \
    python tools/bench_split.py > gpurun_out/ncu_split.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_split.log
