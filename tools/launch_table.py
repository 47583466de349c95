"""Per-kernel totals and per-launch durations from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: launch_table.py launches.csv [kernel_substring ...]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
seq = []
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(',', '')); u = r[ui]
        ms = v / 1e6 if u == 'ns' else v / 1e3 if u in ('us', 'usecond') else v
        seq.append((r[ki].split('(')[0].replace('void ', ''), ms))
tot = collections.defaultdict(float); cnt = collections.Counter()
for n, ms in seq: tot[n] += ms; cnt[n] += 1
all_ms = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{n:40s} launches {cnt[n]:4d}  total {v:9.3f} ms  share {v / all_ms:.3f}")
for name in sys.argv[2:]:
    print(name, [round(ms, 2) for n, ms in seq if name in n])
