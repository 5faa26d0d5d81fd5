"""A/B builds of the library: python tools/ab.py name1="-DX=0 -DY=1" name2="..."  ->  tools/ab/libglg_<name>.so

Only glg_race.cu is recompiled per variant (the other objects are reused from the in-tree build); a bench or test picks
a variant with GLG_LIB_PATH=tools/ab/libglg_<name>.so.  Prints registers / spills of the fused and packed kernels."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from game_level_gan_b200 import build as b
b.build()
out_dir = os.path.join(ROOT, 'tools', 'ab')
os.makedirs(out_dir, exist_ok=True)
procs = []
for arg in sys.argv[1:]:
    name, flags = arg.split('=', 1)
    obj = os.path.join(out_dir, 'glg_race_%s.o' % name)
    cmd = [b.NVCC] + b.FLAGS + flags.split() + ['-Xptxas', '-v', '-c', os.path.join(b.CSRC, 'glg_race.cu'), '-o', obj]
    procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
for name, obj, p in procs:
    log = p.communicate()[0].decode()
    if p.returncode:
        sys.exit('%s: nvcc failed\n%s' % (name, log))
    lines = log.split('\n')
    for i, l in enumerate(lines):
        if 'Compiling entry function' in l and ('race_rollout_fused_kernelILi2ELi130ELi2' in l or 'race_rollout_fused_kernelILi1ELi130ELi4' in l):
            print(name, re.search(r"'(_Z\w+)'", l).group(1)[:60], '|', lines[i + 1].strip(), '|', lines[i + 2].strip())
    objs = [obj] + [os.path.join(b.CSRC, s.replace('.cu', '.o')) for s in b.SOURCES if s != 'glg_race.cu']
    lib = os.path.join(out_dir, 'libglg_%s.so' % name)
    subprocess.check_call([b.NVCC, '-shared', '-o', lib] + objs + ['-lcudart'])
    print(name, '->', os.path.relpath(lib, ROOT))
