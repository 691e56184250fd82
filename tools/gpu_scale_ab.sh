#!/bin/bash
# N-GPU A/B of the gradient all-reduce modes:  N=2 MODES="flat bucket bucket:32" bash tools/gpu_scale_ab.sh   (mode:SMs)
N=${N:-2}
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
for m in ${MODES:-bucket flat}; do
  mode=${m%%:*}; sms=32; [[ "$m" == *:* ]] && sms=${m##*:}
  EDG_ALLREDUCE_SMS=$sms timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $N --steps 50 --warmup 10 --no-cpu-baseline --allreduce $mode > gpurun_out/scale_${mode}_${N}.json 2> gpurun_out/scale_${mode}_${N}.err
  echo "mode=$m N=$N rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${mode}_${N}.json').read().strip().splitlines()[-1])
    print('ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']))
except Exception as e:
    print('no line', e)
PY
)"
  grep -E "n_late|Error|error" gpurun_out/scale_${mode}_${N}.err | head -5
done
done
