#!/bin/bash
# whole GPU suite, bench (with cpu baseline), reference arm, ncu launch list, config-5 sweep
bash tools/gpu_tests.sh
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_full.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python tools/bench_c5.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "c5 rc=$?"; cat gpurun_out/bench_c5.json; tail -3 gpurun_out/bench_c5.err
