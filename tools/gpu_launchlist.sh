#!/bin/bash
# ncu launch list of a short bench run (the graph replays are profiled node by node; cache state kept): isolated
# duration and DRAM bytes of every launch.  STEPS timed steps after the per-kernel profiling pass.
mkdir -p gpurun_out
STEPS=${STEPS:-3}
python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 3000 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -c 300 gpurun_out/plain.log; wc -l gpurun_out/launches.csv
