"""Print the essentials of gpurun_out/bench.json (last JSON line): step time, throughput, the slowest kernels."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/bench.json"
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
except Exception as e:
    print(tag, "no bench line:", e)
    print(open("gpurun_out/bench.err").read()[-1500:])
    sys.exit(0)
print(tag, "ms_per_step %.4f  value %.0f  e2e %.0f  launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("gpu_launches")))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"])[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
    print("  %-48s %7.1f x%s" % (k, v["us_per_launch"], v["launches_per_step"]))
