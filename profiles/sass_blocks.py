"""Instruction and stall-sample distribution of an `ncu --page source --csv` export in blocks of N SASS lines.
usage: python profiles/sass_blocks.py file.csv units_per_launch [block=50] [min_per_unit=30]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); blk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
mn = float(sys.argv[4]) if len(sys.argv) > 4 else 30
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
i_src, i_s, i_ex, i_t = (hdr.index(k) for k in ('Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed'))
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(i_src, i_s, i_ex) and r[i_s].isdigit()]
tot = sum(int(r[i_ex]) for r in data); ts = sum(int(r[i_s]) for r in data)
print('warp-instr per unit', round(tot / units, 1), 'samples', ts)
for a in range(0, len(data), blk):
    e = sum(int(r[i_ex]) for r in data[a:a + blk]); s = sum(int(r[i_s]) for r in data[a:a + blk])
    t = sum(int(r[i_t]) for r in data[a:a + blk])
    if e / units > mn:
        print(f"{a:5d} instr/unit {e/units:8.1f} ({100*e/tot:4.1f}%) samples {100*s/ts:4.1f}% lanes {t/max(e,1):5.1f}  {data[a][i_src][:56]}")
