"""Per-CTA timeline of the last step launch (build with GLG_NVCC_EXTRA=-DGLG_PHASE_CLOCKS)."""
import ctypes, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200 import _lib
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
reps = [bench.Replica(i, 0, dev, 'fast') for i in range(4)]
h = ctypes.CDLL(_lib.LIB_PATH)
for i in range(8): reps[i % 4].cycle(100)
n = 4096 * 4
buf = (ctypes.c_ulonglong * n)()
h.glg_debug_trace(buf, n)
a = np.array(buf[:], dtype=np.int64).reshape(4096, 4)
t0 = a[:, 0].min()
start, wait, end, sm = a[:, 0] - t0, a[:, 1] - t0, a[:, 2] - t0, a[:, 3]
print('launch span [us]: first start 0, last start %.2f, first end %.2f, last end %.2f' % (start.max() / 1e3, end.min() / 1e3, end.max() / 1e3))
print('wait release [us]: min %.2f max %.2f' % (wait.min() / 1e3, wait.max() / 1e3))
dur = (end - wait) / 1e3
print('active lifetime after wait [us]: mean %.2f p10 %.2f p50 %.2f p90 %.2f max %.2f' % (dur.mean(), *np.percentile(dur, [10, 50, 90]), dur.max()))
print('resident time incl. wait [us]: mean %.2f' % ((end - start).mean() / 1e3))
# concurrency over time (active = after wait)
ts = np.arange(0, end.max(), 500)
for t in ts:
    act = int(((wait <= t) & (end > t)).sum()); res = int(((start <= t) & (end > t)).sum())
    print('t=%6.2f us  active CTAs %5d  resident (this launch) %5d' % (t / 1e3, act, res))
order = np.argsort(wait)
print('CTA index vs release time: first released', order[:8], 'last', order[-8:])
print('lifetime of early vs late CTAs: first 2368 mean %.2f, rest mean %.2f' % (dur[order[:2368]].mean(), dur[order[2368:]].mean()))
