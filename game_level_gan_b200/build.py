"""Build recipe of the CUDA library (sm_100a only, in-tree so the .so travels with the repo).

    python -m game_level_gan_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  --fmad=false: the bit-exact predicates must not be contracted
into FMAs (csrc/glg_exact.cuh); code that wants an FMA calls fmaf() explicitly.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(CSRC, 'libglg_b200.so')
SOURCES = ['glg_abi.cu', 'glg_track.cu', 'glg_race.cu', 'glg_helpers.cu', 'glg_pacman.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '--fmad=false',
         '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--expt-relaxed-constexpr']


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'glg_b200.h'))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    extra = os.environ.get('GLG_NVCC_EXTRA', '').split()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        cmd = [NVCC] + FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write('---- %s ----\n%s\n' % (src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([NVCC, '-shared', '-o', OUT] + objs + ['-lcudart'])
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
