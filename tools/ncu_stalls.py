import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__cycles_active.avg','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__waves_per_multiprocessor']
for r in rows[2:]:
    print('----')
    for k in want:
        if k in hdr: print(' ',k, r[hdr.index(k)])
    st=[(float(r[i] or 0),h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    for v,h in sorted(st,reverse=True)[:8]: print('  stall %-28s %.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v))
