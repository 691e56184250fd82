#!/bin/bash
# same-box A/B of two builds of libedgcn (EDG_LIB): alternate runs, print ms/step.   usage: gpu_ab.sh <libA> <libB>
mkdir -p gpurun_out
A=${1:-csrc/libedgcn_prev.so}; B=${2:-csrc/libedgcn.so}
for i in 1 2; do
  for lib in $A $B; do
    EDG_LIB=$PWD/ed-gated-gcn_b200/$lib timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
k=d['kernels'].get('edg_head_bwd',{}).get('us_per_launch')
print("$lib", 'ms_per_step', round(d['ms_per_step'],4), 'head_bwd us', k)
PY
  done
done
