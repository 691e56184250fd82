#!/bin/bash
# ncu --set full of the fused layer kernel (micro-benchmark, one configuration from the environment)
mkdir -p gpurun_out
timeout 300 python tools/bench_fused.py quick > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcn_layer -s 6 -c 2 -f -o gpurun_out/prof_fused \
    python tools/bench_fused.py quick > gpurun_out/ncu_fused.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu_fused.log
