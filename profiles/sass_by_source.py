"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g -c <cubin>` line info and report the executed
warp-instructions per SOURCE line.  usage: python profiles/sass_by_source.py ncu_source.csv nvdisasm.txt kernel_substr units [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
dis = open(sys.argv[2]).read().splitlines()
kern, units = sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = [r for r in rows[rows.index(hdr) + 1:] if len(r) > max(i_src, i_s, i_ex) and r[i_s].isdigit()]
# instruction sequence of the kernel in the disassembly with the current source line
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
print('sass in csv', len(data), 'in disassembly', len(seq))
n = min(len(data), len(seq))
agg = collections.Counter(); smp = collections.Counter()
for (ln, _), r in zip(seq[:n], data[:n]):
    agg[ln] += int(r[i_ex]); smp[ln] += int(r[i_s])
tot = sum(agg.values()); ts = sum(smp.values()) or 1
for ln, e in agg.most_common(top):
    print(f"{str(ln):28s} {e/units:9.1f} instr/unit {100*e/tot:5.1f}%  stall {100*smp[ln]/ts:5.1f}%")
