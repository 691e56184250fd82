#!/bin/bash
# smoke, whole GPU suite, bench (what the driver runs at round end)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^(FAILED|ERROR)|^E " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_full.json
