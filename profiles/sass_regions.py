"""Bucket an `ncu --page source --csv` export into contiguous SASS regions of similar execution count.
usage: python profiles/sass_regions.py file.csv units_per_launch [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(i_src, i_s, i_ex) and r[i_s].isdigit()]
# ncu lists every line twice in this export mode; keep the first copy
half = len(data) // 2
if half and all(data[i][i_src] == data[i + half][i_src] for i in range(0, half, max(1, half // 50))):
    data = data[:half]
tot = sum(int(r[i_ex]) for r in data) or 1; ts = sum(int(r[i_s]) for r in data) or 1
print('sass lines', len(data), 'warp-instr', tot, 'per unit', round(tot / units, 1), 'samples', ts)
reg, cur = [], None
for idx, r in enumerate(data):
    e, s = int(r[i_ex]), int(r[i_s])
    if cur and abs(e - cur['e']) <= 0.25 * max(e, cur['e'], 1):
        cur['n'] += 1; cur['sum'] += e; cur['s'] += s; cur['end'] = idx
    else:
        cur = {'start': idx, 'end': idx, 'e': e, 'n': 1, 'sum': e, 's': s}; reg.append(cur)
for g in reg:
    if g['sum'] > min_share * tot or g['s'] > min_share * ts:
        print(f"lines {g['start']:5d}-{g['end']:5d} n={g['n']:4d} exec/line/unit={g['e']/units:8.2f} instr/unit={g['sum']/units:9.1f} "
              f"share={100*g['sum']/tot:5.1f}% stall={100*g['s']/ts:5.1f}%  {data[g['start']][i_src][:50]}")
