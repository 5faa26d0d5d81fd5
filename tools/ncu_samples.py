"""Warp-stall samples (time) and executed instructions per source line of the first kernel in an ncu report."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur = None; nblocks = 0; seen = set(); stop = False
agg = {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        if cur in seen: break
        seen.add(cur); continue
    if r[0] in ('Function Name', 'Line No', 'Kernel Name', 'File Name'): continue
    if len(r) > 8 and r[0].isdigit() and r[2] == '-':
        try: samp = int(r[6]); inst = int(r[7])
        except ValueError: continue
        key = (cur, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0]); a[0] += samp; a[1] += inst
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print('total samples', ts, 'total inst', ti)
for (f, l, s), (samp, inst) in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    print('%5.1f%% time %5.1f%% inst  %s:%d  %s' % (100. * samp / ts, 100. * inst / ti, f, l, s))
