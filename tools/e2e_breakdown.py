import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceConfig
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
tape, snap, _ = bench.record_tape(env, bench.synthetic_tracks(4096, 5), 6, dev)
host_acts = tape.cpu().pin_memory()
out_s = torch.empty((2, 4096, 20)).pin_memory(); out_r = torch.empty((2, 4096)).pin_memory()
def loop(k, mode):
    t = {'h2d': 0., 'step': 0., 'd2h': 0., 'sync': 0.}
    for s in range(k):
        if s % 100 == 0: env.restore(snap)
        t0 = time.perf_counter()
        a = host_acts[100 + s % 100].to(dev, non_blocking=True)
        t1 = time.perf_counter()
        st, rw = env.step(a)
        t2 = time.perf_counter()
        out_s.copy_(st, non_blocking=True); out_r.copy_(rw, non_blocking=True)
        t3 = time.perf_counter()
        if mode == 'sync': torch.cuda.current_stream().synchronize()
        t4 = time.perf_counter()
        t['h2d'] += t1 - t0; t['step'] += t2 - t1; t['d2h'] += t3 - t2; t['sync'] += t4 - t3
    torch.cuda.synchronize()
    return {k_: 1e6 * v / k for k_, v in t.items()}
for mode in ('async', 'sync'):
    loop(50, mode)
    t0 = time.perf_counter(); r = loop(400, mode); dt = time.perf_counter() - t0
    print(mode, 'us/step %.1f' % (1e6 * dt / 400), {k: round(v, 1) for k, v in r.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); loop(400, 'async'); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
