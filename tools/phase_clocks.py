"""Per-phase warp latency of the step kernel (build with GLG_NVCC_EXTRA=-DGLG_PHASE_CLOCKS)."""
import ctypes, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200 import _lib
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
if len(sys.argv) > 1: bench.B_TRACKS = int(sys.argv[1])
rep = bench.Replica(0, 0, dev, 'fast')
h = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 32)()
for _ in range(3): rep.cycle(100)
h.glg_debug_phases(buf, 1)
n = 20
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): rep.cycle(100)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('us per step', 1e6 * dt / (n * 100))
h.glg_debug_phases(buf, 1)
warps = n * 100 * bench.B_TRACKS
names = {0: 'launch, TMA issue, chain/grid wait', 1: 'state loads + kinematics', 2: 'barrier + TMA wait', 3: 'progress arg-min', 4: 'scan pre + ray table', 5: 'stage 1', 6: 'lists', 7: 'collision', 8: 'finish/reward/write-back/release', 9: 'stage 2 + first-ray evaluation', 10: 'second evaluation round', 11: 'pack'}
old_names = {0: 'launch..griddep wait', 1: 'state loads + kinematics', 2: 'syncthreads + TMA wait', 3: 'progress arg-min', 5: 'scan pre + stage 1', 6: 'lists', 7: 'collision', 8: 'stage 2 + emit', 9: 'reward/finish (rest of main between 3 and 9 minus scan)', 10: 'flush (exact eval)', 11: 'state write + (sensors_finish rest)', 12: 'pack'}
tot = sum(buf[i] for i in range(32))
for i in range(32):
    if buf[i]: print('%2d %-55s %8.0f cycles/warp %5.1f%%' % (i, names.get(i, ''), buf[i] / warps, 100. * buf[i] / tot))
print('sum', tot / warps)
