#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"; grep "smoke " gpurun_out/smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|fc_head_fwd|fc_head_bwd|wgrad_tc_batch|linear_ws|aggregate_staged" -s 26 -c 13 -o gpurun_out/final_r1h -f python tools/prof_kernels2.py > gpurun_out/ncu_final_h.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_final_h.log
