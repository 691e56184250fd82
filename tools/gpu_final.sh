#!/bin/bash
# final validation of a round: smoke, whole GPU suite, bench + reference arm, ncu launch list (durations + DRAM bytes),
# ncu --set full of the dominant kernels (micro-benchmarks), config-5 sweep, config-3 line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"
# exactly what the driver runs (one process)
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^(FAILED|ERROR)|^E " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
STEPS=3 bash tools/gpu_launchlist.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_layer -s 6 -c 4 -f -o gpurun_out/prof_fused \
    python tools/bench_fused.py quick > gpurun_out/ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"head_du_kernel|scores_staged" -s 4 -c 4 -f -o gpurun_out/prof_head \
    python tools/bench_head.py > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?"
timeout 300 python tools/bench_fused.py quick > gpurun_out/bench_fused.log 2>&1; tail -8 gpurun_out/bench_fused.log
timeout 300 python tools/bench_head.py > gpurun_out/bench_head.log 2>&1; tail -4 gpurun_out/bench_head.log
timeout 600 python tools/bench_c5.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "c5 rc=$?"; cut -c1-600 gpurun_out/bench_c5.json
timeout 900 python bench.py --config C3 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "c3 rc=$?"; cut -c1-400 gpurun_out/bench_c3.json
