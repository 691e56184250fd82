#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"staged|views_bwd" -s 6 -c 6 -o gpurun_out/staged_r1 -f python tools/prof_kernels.py > gpurun_out/ncu_staged.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_staged.log
