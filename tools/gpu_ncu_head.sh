#!/bin/bash
# ncu --set full of one kernel of tools/bench_head.py:  KERNEL=<regex> WHICH=scores|head bash tools/gpu_ncu_head.sh
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KERNEL:-scores_staged}" -s 6 -c 2 -f -o gpurun_out/prof_head \
    python tools/bench_head.py ${WHICH:-scores} > gpurun_out/ncu_head.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_head.log
