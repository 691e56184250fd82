#!/bin/bash
# light final validation (no ncu): smoke, whole GPU suite, bench + reference arm, the fp32-parity GEMM bench, the drop-in
# lines, config-3 line.  The ncu captures live in tools/gpu_final.sh / gpu_ncu_*.sh / gpu_launchlist.sh.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "^(FAILED|ERROR)|^E " gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; python tools/show_bench.py full gpurun_out/bench_full.json 30
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 python bench.py --no-cpu-baseline --torch-head > gpurun_out/bench_torch_head.json 2> gpurun_out/bench_torch_head.err
python tools/show_bench.py torch-head gpurun_out/bench_torch_head.json 0
timeout 200 python tools/bench_split.py > gpurun_out/bench_split.txt 2>&1; cat gpurun_out/bench_split.txt
for d in f32 bf16; do timeout 200 python tools/bench_dropin.py --dtype $d 2>/dev/null | tail -1 > gpurun_out/bench_dropin_$d.json; cut -c1-420 gpurun_out/bench_dropin_$d.json; done
timeout 900 python bench.py --config C3 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "c3 rc=$?"; python tools/show_bench.py c3 gpurun_out/bench_c3.json 6
