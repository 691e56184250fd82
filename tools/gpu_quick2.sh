#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "$1" > gpurun_out/quick_tests.log 2>&1
echo "== tests exit=$? : $(tail -1 gpurun_out/quick_tests.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/quick_tests.log | head
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(f"  {k:45s} {v['us_per_launch']:8.1f} x{v['launches_per_step']}")
PY
EDG_AGG_VARIANT=0 timeout 300 python tools/bench_c5.py --chunks 4 --cpu-graphs 256 2>&1 | cut -c1-700
