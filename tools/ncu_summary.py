"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep stays in gpurun_out/).
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep  > profiles/rNN_ncu_full.md
"""
import collections, csv, subprocess, sys

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict(); n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        if "spin_kernel" in row["Kernel Name"]:       # bench.py's preload of the per-kernel pass, not part of a step
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0]); a[0] += 1; a[1] += v; n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"ncu launch list ({path}): {n} launches, {tot:.0f} us of kernel time (cold-cache, serialised: compare shares)\n")
    print("| share | launches | avg us | kernel |\n|---:|---:|---:|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.002:
            continue
        print(f"| {a[1] / tot * 100:.1f}% | {a[0]} | {a[1] / a[0]:.1f} | `{k}` |")

WANT = [("gpu__time_duration.sum", "dur us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst")]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"ncu --set full --clock-control none ({path}); one row per captured launch\n")
    print("| kernel | " + " | ".join(n for _, n in WANT) + " |\n|---|" + "---:|" * len(WANT))
    for d in data:
        cells = []
        for m, _ in WANT:
            if m in hdr:
                v = d[hdr.index(m)].replace(",", "")
                try:
                    f = float(v)
                    u = units[hdr.index(m)]
                    if m.startswith("dram__bytes"):
                        f = f / 1e6 if u == "byte" else f * 1e3 if u == "Gbyte" else f / 1e3 if u == "Kbyte" else f
                    if m == "gpu__time_duration.sum":
                        f = f / 1e3 if u == "ns" else f * 1e3 if u == "ms" else f
                    cells.append(f"{f:.1f}" if f < 1e6 else f"{f:.3g}")
                except ValueError:
                    cells.append(v)
            else:
                cells.append("-")
        print(f"| `{d[hdr.index('Kernel Name')][:60]}` | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
