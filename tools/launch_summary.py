"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and isolated
duration over the LAST n launches (default: everything after the last spin kernel = the final graph replays)."""
import csv, collections, re, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value'); mn = hdr.index('Metric Name')
seq = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv and r[mn] == 'gpu__time_duration.sum']
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
last_spin = max((i for i, (n, _) in enumerate(seq) if 'spin_kernel' in n), default=-1)
tail = seq[last_spin + 1:]
def short(n):
    n = re.sub(r'\(.*', '', n)
    n = re.sub(r'^void ', '', n)
    return n[-70:]
agg = collections.OrderedDict()
for n, v in tail:
    agg.setdefault(short(n), []).append(v)
tot = sum(v for _, v in tail)
print(f"{len(tail)} launches after the profiling pass, {tot / 1e3:.1f} us serialised ({tot / 1e3 / steps:.1f} us per step over {steps} steps)")
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-72s n=%3d avg=%8.1f us  per-step=%8.1f us' % (n, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3 / steps))
