#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 3; do
  echo "== EDG_FUSED_DEBUG=$dbg (1: phase B skipped, 2: phase A skipped)"
  EDG_FUSED_DEBUG=$dbg EDG_FUSED_NPASS=1 timeout 300 python tools/bench_fused.py quick 2>&1 | grep fused
done
