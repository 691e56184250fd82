#!/bin/bash
# quick loop: the fused-kernel tests, then the bench line with its per-kernel table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f_fused.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/fused_tests.log 2>&1
echo "== tests exit=$? : $(tail -1 gpurun_out/fused_tests.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/fused_tests.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'],'value',d['value'],'e2e',d.get('e2e',{}).get('value'))
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:24]: print('  %-48s %7.1f x%s'%(k,v['us_per_launch'],v['launches_per_step']))
PY
tail -3 gpurun_out/bench.err
