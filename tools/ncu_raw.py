"""Key raw metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` dump (development tool). usage: ncu_raw.py X.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
want += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for vals in rows[2:]:
    for h, u, v in zip(hdr, rows[1], vals):
        if h in want and v not in ('0', '0.000000'):
            print(f"{h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '')} [{u}] {v}")
    print('---')
