#!/bin/bash
# bench line, then (only if it exited 0) the ncu launch list of the same command line
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?
echo "bench rc=$rc"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ $rc -eq 0 ] && [ "$1" == "ncu" ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
fi
