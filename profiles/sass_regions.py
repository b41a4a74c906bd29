"""Executed warp-instructions and stall samples of one kernel per SOURCE-LINE REGION (ncu source page joined with
nvdisasm line info, like sass_by_source.py).
usage: python profiles/sass_regions.py ncu_source.csv nvdisasm.txt kernel_substr units file name:lo-hi [name:lo-hi ...]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
dis = open(sys.argv[2]).read().splitlines()
kern, units, fname = sys.argv[3], float(sys.argv[4]), sys.argv[5]
regions = []
for a in sys.argv[6:]:
    n, r = a.split(':'); lo, hi = r.split('-'); regions.append((n, int(lo), int(hi)))
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(i_src, i_s, i_ex) and r[i_s].isdigit()]
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l)
seq, line = [], None
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if seq:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((line, m.group(2).strip()))
n = len(seq)
agg = collections.Counter(); smp = collections.Counter(); nsass = collections.Counter()
last = 'other'
for (ln, _), r in zip(seq, data[:n]):
    name = None
    if ln and ln[0] == fname:
        for rn, lo, hi in regions:
            if lo <= ln[1] <= hi:
                name = rn
                break
        last = name or 'other'
    else:
        name = last          # inlined helpers (common.cuh, device headers) go to the region that called them
    agg[name or 'other'] += int(r[i_ex]); smp[name or 'other'] += int(r[i_s]); nsass[name or 'other'] += 1
tot = sum(agg.values()); ts = sum(smp.values()) or 1
print('sass instructions', n, ' executed warp-instr per unit %.0f' % (tot / units))
for rn in [r[0] for r in regions] + ['other']:
    print(f"{rn:22s} {nsass[rn]:5d} sass {agg[rn]/units:9.1f} instr/unit {100*agg[rn]/tot:5.1f}%  stall samples {100*smp[rn]/ts:5.1f}%")
