"""Print the headline numbers of bench.py JSON lines:  python tools/show_bench.py gpurun_out/*.json"""
import json, sys
for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith('{')]
    if not lines:
        print('==', path, '(no JSON line)'); continue
    d = json.loads(lines[0])
    print('==', path)
    print('  value %.4g  ms/step %.5f  frac %.4f  e2e %.4g  launches %s  repeats %s  n_gpus %s' % (
        d['value'], d['ms_per_step'], d.get('roofline', {}).get('frac', 0), d['e2e']['value'], d.get('gpu_launches'),
        d['config'].get('repeats'), d.get('n_gpus')))
    if d['config'].get('block_ms'): print('  block_ms', [round(x, 4) for x in d['config']['block_ms']])
    print('  clocks', d.get('clocks'), ' parity', d.get('parity'))
    for k, v in (d.get('extra') or {}).items(): print('  extra', k, json.dumps(v)[:700])
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'])
    if 'sanity' in d: print('  SANITY', d['sanity'])
