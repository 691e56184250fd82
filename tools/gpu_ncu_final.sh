#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"linear_ws|wgrad_tall|aggregate_staged|pool_staged|scores_staged|head_bwd_staged|views_bwd|split_reduce" -s 9 -c 9 -o gpurun_out/final_r1 -f python tools/prof_kernels.py > gpurun_out/ncu_final.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_final.log
