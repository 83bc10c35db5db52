"""Selected columns of an ncu raw page (csv on stdin or file) -> compact csv: python scripts/ncu_summary.py raw.csv out.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__cycles_active.sum', 'sm__cycles_elapsed.max', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
idx = [hdr.index(k) for k in keep if k in hdr]
w = csv.writer(open(sys.argv[2], 'w', newline=''))
w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
