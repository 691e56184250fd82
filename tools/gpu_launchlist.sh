#!/bin/bash
# ncu launch list of a short bench run (graph replay is profiled node by node); cache state kept
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -c 400 gpurun_out/plain.log; wc -l gpurun_out/launches.csv
