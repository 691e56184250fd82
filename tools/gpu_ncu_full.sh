#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"aggregate_kernel|linear_tc|wgrad_tc|head_bwd|pool_fwd|scores_kl|views_bwd|colsum_partial" -s 9 -c 9 -o gpurun_out/prof_r1 -f python tools/prof_kernels.py > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
