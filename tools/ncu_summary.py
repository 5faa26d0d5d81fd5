"""Summarise an ncu --set full raw CSV export into a small markdown table (for profiles/)."""
import csv, sys
raw, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
with open(out, 'w') as f:
    f.write('# %s\n\n| metric | unit | %s |\n|---|---|%s\n' % (title, ' | '.join('launch %d' % i for i in range(len(data))), '---|' * len(data)))
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write('| %s | %s | %s |\n' % (w, units[i], ' | '.join(r[i] for r in data)))
print(open(out).read())
