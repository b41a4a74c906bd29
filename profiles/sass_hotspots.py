"""Summarise an `ncu --page source --csv` export: stall samples and executed instructions per opcode
and the hottest SASS lines.  usage: python profiles/sass_hotspots.py file.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(i_src, i_s, i_ex) and r[i_s].isdigit()]
tot = sum(int(r[i_s]) for r in data) or 1
totex = sum(int(r[i_ex]) for r in data) or 1
print('sass lines', len(data), 'stall samples', tot, 'warp instructions executed', totex)
op, opx = collections.Counter(), collections.Counter()
for r in data:
    toks = r[i_src].split()
    o = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op[o] += int(r[i_s]); opx[o] += int(r[i_ex])
for o, c in op.most_common(top):
    print(f"{o:28s} samples {c:8d} ({100*c/tot:5.1f}%)  executed {opx[o]:11d} ({100*opx[o]/totex:5.1f}%)")
print('--- hottest lines')
for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][i_s]))[:top]:
    print(f"{idx:5d} samples {r[i_s]:>7s} exec {r[i_ex]:>9s}  {r[i_src][:100]}")
