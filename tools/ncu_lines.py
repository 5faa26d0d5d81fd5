"""Per-source-line executed warp instructions of the first kernel instance in an ncu report."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60; cars = float(sys.argv[3]) if len(sys.argv) > 3 else 8192.
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; blocks = []; agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; blocks.append(cur); continue
    if r[0] in ('Function Name', 'Line No'): continue
    if len(r) > 8 and r[2] == '-' and r[0].isdigit():
        key = (len(blocks), cur, int(r[0]), r[1].strip()[:95])
        try: agg[key] = agg.get(key, 0) + int(r[7])
        except ValueError: pass
first = []; names = set()
for n, f in enumerate(blocks, 1):
    if f in names: break
    names.add(f); first.append(n)
agg = {k: v for k, v in agg.items() if k[0] in first}
# the per-line numbers double count inlined callees; report per-car values of leaf lines
tot = sum(agg.values())
byfile = {}
for (b, f, l, s), v in agg.items(): byfile[f] = byfile.get(f, 0) + v
print('sum over lines', tot, {k: round(v / cars, 1) for k, v in byfile.items()})
for (b, f, l, s), v in sorted(agg.items(), key=lambda x: -x[1])[:topn]:
    print('%7.1f /car  %s:%d  %s' % (v / cars, f, l, s))
