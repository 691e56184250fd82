#!/bin/bash
# same-box A/B of two builds of libedgcn (EDG_LIB): alternate runs, print ms/step
mkdir -p gpurun_out
for i in 1 2; do
  for lib in csrc/libedgcn_prev.so csrc/libedgcn.so; do
    EDG_LIB=$PWD/ed-gated-gcn_b200/$lib timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print("$lib", 'ms_per_step', round(d['ms_per_step'],4))
PY
  done
done
