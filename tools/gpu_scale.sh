#!/bin/bash
# the driver's N-GPU bench launch:  N=8 bash tools/gpu_scale.sh   -> gpurun_out/bench_${N}gpu.json
N=${N:-2}
mkdir -p gpurun_out
[ -n "$SMOKE" ] && { timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? : $(tail -1 gpurun_out/smoke.log)"; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "N=$N rc=$?"; python tools/show_bench.py "N=$N" gpurun_out/bench_${N}gpu.json 0
grep -E "Error|error" gpurun_out/bench_${N}gpu.err | head -5
