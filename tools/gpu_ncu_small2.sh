#!/bin/bash
mkdir -p gpurun_out
python tools/prof_small2.py > gpurun_out/prof_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fc_head|mlp_chain|wgrad_tc_batch" -s 6 -c 6 -o gpurun_out/small2 -f python tools/prof_small2.py > gpurun_out/ncu_small2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/prof_small2.log gpurun_out/ncu_small2.log
