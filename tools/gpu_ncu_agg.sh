#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do
  EDG_AGG_VARIANT=$v python tools/prof_agg.py > gpurun_out/agg_plain$v.log 2>&1 &&
  EDG_AGG_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:"aggregate" -s 2 -c 2 -o gpurun_out/agg_v$v -f python tools/prof_agg.py > gpurun_out/ncu_agg$v.log 2>&1
  echo "v$v rc=$?"
done
ls -la gpurun_out/agg_v*.ncu-rep
