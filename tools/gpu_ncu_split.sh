#!/bin/bash
# ncu --set full of the fp32-parity (split operand) kernels of tools/bench_split.py, then the drop-in bench lines
mkdir -p gpurun_out
timeout 200 python tools/bench_split.py > gpurun_out/bench_split.txt 2>&1; echo "bench_split rc=$?"; cat gpurun_out/bench_split.txt
for d in f32 bf16; do timeout 200 python tools/bench_dropin.py --dtype $d 2>/dev/null | tail -1 > gpurun_out/bench_dropin_$d.json; cat gpurun_out/bench_dropin_$d.json; done
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"linear_ws_kernel|wgrad_tall_kernel|split_f16_kernel|amax_kernel|split_reduce_bias_scaled" -c 14 -f -o gpurun_out/prof_split \
    python tools/bench_split.py > gpurun_out/ncu_split.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_split.log
