#!/bin/bash
# final validation of the round: smoke, whole GPU suite, bench + reference arm, ncu launch list, ncu --set full of the
# kernels that changed this session (one step outside CUDA graphs), config-5 sweep
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"; grep smoke gpurun_out/smoke.log
bash tools/gpu_tests.sh
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
timeout 300 python tools/prof_kernels2.py > gpurun_out/prof2_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|fc_head_fwd|fc_head_bwd|wgrad_tc_batch|linear_ws|aggregate_staged" -s 40 -c 14 -o gpurun_out/final_r1h -f python tools/prof_kernels2.py > gpurun_out/ncu_final_h.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_final_h.log
timeout 600 python tools/bench_c5.py > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
echo "c5 rc=$?"; cut -c1-900 gpurun_out/bench_c5.json
