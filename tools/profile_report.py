"""profiles/<name>.md from an ncu report: headline metrics, warp stall breakdown, hottest source lines.

    python tools/profile_report.py gpurun_out/prof_X.ncu-rep profiles/X_step_kernel.md "title" [cars]
"""
import csv, subprocess, sys, io
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
cars = float(sys.argv[4]) if len(sys.argv) > 4 else 8192.
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
L = ['# %s' % title, '', '| metric | unit | %s |' % ' | '.join('launch %d' % i for i in range(len(data))), '|---|---|' + '---|' * len(data)]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        L.append('| %s | %s | %s |' % (w, units[i], ' | '.join(r[i] for r in data)))
i = hdr.index('smsp__inst_executed.sum')
L += ['', 'warp instructions per car (launch 0): %.0f' % (float(data[0][i]) / cars), '', '## warp stall reasons (launch 0, smsp__average_warps_issue_stalled_*_per_issue_active.ratio)', '']
st = []
for j, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
        try: st.append((float(data[0][j]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        except ValueError: pass
for v, n in sorted(st, reverse=True)[:10]:
    L.append('* %-22s %.3f' % (n, v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur, blocks, agg = None, [], {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; blocks.append(cur); continue
    if r[0] in ('Function Name', 'Line No'): continue
    if len(r) > 8 and r[2] == '-' and r[0].isdigit():
        key = (len(blocks), cur, int(r[0]), r[1].strip()[:100])
        try: agg[key] = agg.get(key, 0) + int(r[7])
        except ValueError: pass
first, names = [], set()
for n, f in enumerate(blocks, 1):
    if f in names: break
    names.add(f); first.append(n)
agg = {k: v for k, v in agg.items() if k[0] in first}
byfile = {}
for (b, f, l, s), v in agg.items(): byfile[f] = byfile.get(f, 0) + v
L += ['', '## executed warp instructions per car by source file (launch 0; inlined callees are attributed to their own file)', '']
for f, v in sorted(byfile.items(), key=lambda x: -x[1]): L.append('* %-32s %.1f' % (f, v / cars))
L += ['', '## hottest source lines (warp instructions per car)', '', '```']
for (b, f, l, s), v in sorted(agg.items(), key=lambda x: -x[1])[:40]:
    L.append('%7.1f  %s:%d  %s' % (v / cars, f, l, s))
L.append('```')
open(out, 'w').write('\n'.join(L) + '\n')
print('\n'.join(L[:60]))
