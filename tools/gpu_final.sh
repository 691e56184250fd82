#!/bin/bash
# final validation of the round: smoke, whole GPU suite, bench + reference arm, ncu launch list, ncu --set full of the
# kernels that changed this session (one step outside CUDA graphs), config-5 sweep
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"; grep smoke gpurun_out/smoke.log
# exactly what the driver runs (one process)
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^(FAILED|ERROR)|^E " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
timeout 600 python tools/bench_c5.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "c5 rc=$?"; cut -c1-900 gpurun_out/bench_c5.json
