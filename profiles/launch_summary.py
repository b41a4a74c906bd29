"""Summarise an ncu launch-list CSV (gpu__time_duration.sum [+ smsp__inst_executed.sum]) per kernel.
usage: python profiles/launch_summary.py launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name'][:44]
    v = float(row['Metric Value'].replace(',', ''))
    u, mname = row['Metric Unit'], row['Metric Name']
    if mname.startswith('gpu__time'):
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    agg.setdefault(name, {}).setdefault(mname, []).append(v)
tot = sum(sum(m.get('gpu__time_duration.sum', [0])) / max(len(m.get('gpu__time_duration.sum', [1])), 1) for m in agg.values())
print(f"{'kernel':46s} {'n':>4s} {'mean us':>10s} {'share':>7s} {'warp-instr':>14s}")
for k, m in agg.items():
    t = m.get('gpu__time_duration.sum', [0]); i = m.get('smsp__inst_executed.sum')
    mean = sum(t) / len(t)
    print(f"{k:46s} {len(t):4d} {mean:10.1f} {100*mean/tot:6.1f}% {(sum(i)/len(i) if i else 0):14.0f}")
