"""From an `ncu --page raw --csv` export of one bench.py run: the per-launch DRAM traffic of the step's logits pass
(profiles/filter_kernel_traffic.json) and the matcher's pipe / memory figures (profiles/matcher_ncu.json) that bench.py
quotes in `roofline.traffic` and `detail.matcher_roofline`.
usage: python profiles/make_ncu_json.py X_raw.csv batch priors classes source-note"""
import csv, json, os, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr, data = rows[0], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
batch, priors, classes, note = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
here = os.path.dirname(os.path.abspath(__file__))


def val(r, k):
    try:
        return float(r[idx[k]].replace(',', ''))
    except Exception:
        return None


def first(sub):
    return next((r for r in data if sub in r[idx['Kernel Name']]), None)


def unit_scale(k, units=rows[1]):
    u = units[idx[k]]
    return {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}.get(u, 1.0)


f = first('filter_kernel')
if f:
    rd = val(f, 'dram__bytes_read.sum') * unit_scale('dram__bytes_read.sum')
    wr = val(f, 'dram__bytes_write.sum') * unit_scale('dram__bytes_write.sum')
    json.dump({"kernel": f[idx['Kernel Name']], "batch": batch, "priors": priors, "classes": classes,
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "duration_us_under_ncu": val(f, 'gpu__time_duration.sum'), "source": note},
              open(os.path.join(here, 'filter_kernel_traffic.json'), 'w'), indent=1)
out = {"source": note, "batch": batch, "priors": priors}
for name, sub in (('search_kernel', 'search_kernel'), ('match_kernel', 'match_kernel')):
    r = first(sub)
    if not r:
        continue
    out[name] = {k2: val(r, k) for k, k2 in (
        ('gpu__time_duration.sum', 'duration_us_under_ncu'), ('smsp__inst_executed.sum', 'warp_instructions'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pipe_pct_of_peak'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64_cycles_active_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_slot_pct'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'), ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_peak'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall_long_scoreboard_per_issue'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall_barrier_per_issue'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'))}
    out[name]['dram_bytes'] = (val(r, 'dram__bytes_read.sum') * unit_scale('dram__bytes_read.sum') +
                               val(r, 'dram__bytes_write.sum') * unit_scale('dram__bytes_write.sum'))
if 'search_kernel' in out:
    out['fp64_pipe_pct_of_peak'] = {"search_kernel": out['search_kernel']['fp64_pipe_pct_of_peak'],
                                    "match_kernel": out.get('match_kernel', {}).get('fp64_pipe_pct_of_peak')}
    json.dump(out, open(os.path.join(here, 'matcher_ncu.json'), 'w'), indent=1)
print('written')
