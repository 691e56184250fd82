"""Join an ncu report's per-instruction stall samples with source lines (no GUI needed).

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name-regex> <mangled-symbol> [top=40] [launch-index=0]

Uses `ncu --page source --csv` (SASS view: samples / instructions per instruction) and `nvdisasm -g` of the cubin
extracted from the in-tree libedgcn.so (instruction offset -> file:line), so the library must be the build that
was profiled (compile with -lineinfo)."""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kre, sym = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ed_gated_gcn_b200", "csrc", "libedgcn.so")
if not os.path.exists(so):
    so = os.path.join(ROOT, "ed-gated-gcn_b200", "csrc", "libedgcn.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text." + sym + ","))
end = next(i for i in range(start + 1, len(dis)) if dis[i].startswith("//-----"))
cur, off2line = None, {}
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        # attribute inlined helpers (mbar_wait, lds_v4, ...) to the line of the kernel that calls them
        if m.group(3):
            cur = (os.path.basename(m.group(3)), int(m.group(4)))
        else:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
# the report may hold several launches: each starts with a "Kernel Name" row
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
a0 = starts[which]
a1 = starts[which + 1] if which + 1 < len(starts) else len(rows)
hdr, data = rows[a0 + 1], [r for r in rows[a0 + 2:a1] if len(r) > 10]
ia, ii, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ia], 16)
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
tot_i = tot_s = 0.0
for r in data:
    line, _ = off2line.get(int(r[ia], 16) - base, (None, ""))
    ie, sm = float(r[ii] or 0), float(r[iS] or 0)
    a = agg[line]
    a[0] += sm; a[1] += ie
    for c in stall_cols:
        v = float(r[c] or 0)
        if v:
            a[2][hdr[c]] += v
    tot_i += ie; tot_s += sm
cache = {}


def text(f, ln):
    if f not in cache:
        for d in ("ed_gated_gcn_b200/csrc", "ed-gated-gcn_b200/csrc"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                cache[f] = open(p).read().split("\n")
                break
        else:
            cache[f] = []
    return cache[f][ln - 1].strip()[:90] if 0 < ln <= len(cache[f]) else ""


print(f"{sym}: {int(tot_i)} warp instructions, {int(tot_s)} samples")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = line if line else ("?", 0)
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{a[0] / tot_s * 100:5.1f}% smp {a[1] / tot_i * 100:5.1f}% inst {f}:{ln:<4d} {text(f, ln)} | "
          + " ".join(f"{k[6:]}={int(v)}" for k, v in st))
