#!/bin/bash
mkdir -p gpurun_out
for pv in 1 0; do
  EDG_VIEWS_PATCH=$pv timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pv$pv.json 2> gpurun_out/bench_pv$pv.err
  echo "views_patch=$pv rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pv$pv.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'loss', d['config'].get('loss'))
PY
done
EDG_VIEWS_PATCH=1 timeout 900 python -m pytest tests/test_gpu_c_models.py tests/test_gpu_d_fullsize.py tests/test_gpu_b_kernels.py -k "not linear and not wgrad" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/quick_tests.log 2>&1
echo "== tests(patch=1) exit=$? : $(tail -1 gpurun_out/quick_tests.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/quick_tests.log | head
EDG_VIEWS_PATCH=1 timeout 300 python tools/determinism.py 2>&1 | grep -v Warn | tail -8
