"""Print a few headline metrics of an ncu raw page (csv): python scripts/ncu_metrics.py file.csv [row]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want:
            print(f'  {h} [{units[i]}] {r[i]}')
        elif h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
            try:
                if float(r[i]) >= 0.4: print(f'  stall {h[34:-23]} {float(r[i]):.2f}')
            except ValueError:
                pass
