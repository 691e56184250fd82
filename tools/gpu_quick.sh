#!/bin/bash
# quick loop: selected tests (args: pytest -k expression), bench, ncu launch list of a short bench run
mkdir -p gpurun_out
if [ -n "$1" ]; then
  timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -k "$1" > gpurun_out/quick_tests.log 2>&1
  echo "== tests exit=$? : $(tail -1 gpurun_out/quick_tests.log)"; grep -E "^(FAILED|ERROR)|Error|assert " gpurun_out/quick_tests.log | head
fi
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(f"  {k:45s} {v['us_per_launch']:8.1f} x{v['launches_per_step']}")
PY
tail -3 gpurun_out/bench.err
if [ "$2" == "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
fi
