"""Digest of an `ncu --page raw --csv` export: one block per distinct kernel with duration, instruction count, pipe /
DRAM / L2 / occupancy / issue figures and the top stall reasons (warps stalled per issued instruction).
usage: ncu -i X.ncu-rep --page raw --csv > X_raw.csv; python profiles/ncu_digest.py X_raw.csv [--all-launches]"""
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('gpu__time_duration.sum', 'duration'), ('smsp__inst_executed.sum', 'warp_inst'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_pipe_pct_of_peak'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64_cycles_active_pct'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma_pipe_pct'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu_pipe_pct'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu_pipe_pct'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu_pipe_pct'),
        ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct_of_peak'),
        ('lts__t_sector_hit_rate.pct', 'l2_hit_pct'), ('l1tex__t_sector_hit_rate.pct', 'l1_hit_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_slot_pct'),
        ('launch__registers_per_thread', 'registers'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dyn_smem'), ('launch__occupancy_limit_shared_mem', 'occ_limit_smem'),
        ('launch__occupancy_limit_registers', 'occ_limit_regs')]
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
seen = set()
for r in data:
    name = r[idx['Kernel Name']]
    if name in seen and '--all-launches' not in sys.argv:
        continue
    seen.add(name)
    print(name)
    for k, v in cols:
        if k in idx:
            print('    %-26s %16s %s' % (v, r[idx[k]], units[idx[k]]))
    st = sorted(((float(r[idx[h]] or 0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                 for h in stalls), reverse=True)[:8]
    print('    stalls (warps per issue):  ' + '  '.join('%s=%.2f' % (n, v) for v, n in st))
