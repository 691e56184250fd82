#!/bin/bash
# plain timing, then ncu --set full of head_du and the scores kernel (micro-benchmark)
mkdir -p gpurun_out
timeout 300 python tools/bench_head.py > gpurun_out/head_plain.log 2>&1; cat gpurun_out/head_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"head_du_kernel|scores_staged" -s 10 -c 4 -f -o gpurun_out/prof_head \
    python tools/bench_head.py > gpurun_out/ncu_head.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_head.log
