import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceConfig
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
tape, snap, _ = bench.record_tape(env, bench.synthetic_tracks(4096, 5), 6, dev)
host_acts = tape.cpu().pin_memory()
hs = env.host_stepper()
env.restore(snap)
T = {'copy': 0., 'replay': 0., 'sync': 0., 'post': 0.}
n = 300
for s in range(n + 20):
    if s % 100 == 0: env.restore(snap)
    a = host_acts[100 + s % 100]
    t0 = time.perf_counter()
    hs.actions_h.copy_(a); hs._base_np[0], hs._base_np[1] = env.steps, env._seq
    t1 = time.perf_counter()
    hs.graph.replay()
    t2 = time.perf_counter()
    torch.cuda.current_stream(dev).synchronize()
    t3 = time.perf_counter()
    env.steps += 1; env._seq += 1; env._alive_known = bool((hs._stamp_np == env._seq).any())
    t4 = time.perf_counter()
    if s >= 20:
        T['copy'] += t1 - t0; T['replay'] += t2 - t1; T['sync'] += t3 - t2; T['post'] += t4 - t3
print({k: round(1e6 * v / n, 1) for k, v in T.items()})
# GPU-side duration of one graph replay
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
tot = 0.
for s in range(50):
    e0.record(); hs.graph.replay(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
print('graph replay GPU time us', 1e3 * tot / 50)
# the pieces alone
x_h = hs.out_h; x_d = hs.out_d
for name, fn in (('D2H 692KB', lambda: x_h.copy_(x_d, non_blocking=True)), ('H2D 64KB', lambda: hs.in_d.copy_(hs.in_h, non_blocking=True))):
    tot = 0.
    for s in range(50):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    print(name, 'us', 1e3 * tot / 50)
