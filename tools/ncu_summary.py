"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep stays in gpurun_out/).
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep  > profiles/rNN_ncu_full.md
"""
import collections, csv, subprocess, sys

def _bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def launches(path, steps=None):
    """Per kernel: launches, isolated duration, DRAM bytes (when the list was taken with dram__bytes_*).  With `steps`,
    only the launches after the last spin kernel (= the graph replays of the timed steps) are counted and per-step totals
    are printed."""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ids = [int(r["ID"]) for r in rows if "spin_kernel" in r["Kernel Name"]]
    first = (max(ids) + 1) if (steps and ids) else 0
    agg = collections.OrderedDict(); seen = set()
    for row in rows:
        if "spin_kernel" in row["Kernel Name"] or int(row["ID"]) < first:   # bench.py's preload of the per-kernel pass
            continue
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0, 0.0])
        v = float(row["Metric Value"].replace(",", ""))
        if row.get("Metric Name") == "gpu__time_duration.sum":
            a[1] += v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
            a[0] += 1
        elif row.get("Metric Name", "").startswith("dram__bytes"):
            a[2] += _bytes(v, row["Metric Unit"])
    n = sum(a[0] for a in agg.values())
    tot = sum(a[1] for a in agg.values())
    dram = sum(a[2] for a in agg.values())
    print(f"ncu launch list ({path}): {n} launches, {tot:.0f} us of kernel time (cold-cache, serialised: compare shares)")
    if steps:
        print(f"per step ({steps} steps): {n / steps:.0f} launches, {tot / steps:.0f} us serialised, "
              f"{dram / steps / 1e6:.0f} MB of DRAM traffic (read + write)")
    print("\n| share | launches | avg us | DRAM MB / launch | kernel |\n|---:|---:|---:|---:|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if a[1] / tot < 0.002:
            continue
        print(f"| {a[1] / tot * 100:.1f}% | {a[0]} | {a[1] / a[0]:.1f} | {a[2] / a[0] / 1e6:.1f} | `{k}` |")

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
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    else:
        full(sys.argv[2])
