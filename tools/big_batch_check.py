"""Config 4 at its full size on ONE GPU: 2^20 tracks x 4 cars - reset, 20 steps, chained rollout, winners; the first
512 tracks are compared with a 512-track environment (same tracks, same actions)."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceCar
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)]]
B, T = 1 << 20, 20
g = torch.Generator().manual_seed(7)
pool = bench.synthetic_tracks(4096, 11)
idx = torch.randint(0, 4096, (B,), generator=g); idx[:512] = torch.arange(512)
tracks = pool[idx].to(dev)
acts_pool = torch.randint(0, 9, (T, 4, 4096), generator=g)
acts_pool = torch.where(torch.rand((T, 4, 4096), generator=g) < 0.6, torch.ones_like(acts_pool), acts_pool)
acts = acts_pool[:, :, idx].to(dev)
big = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, device=dev)
small = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
sb, _ = big.reset(tracks)
torch.cuda.synchronize(); print('reset of %d tracks: %.1f ms' % (B, 1e3 * (time.perf_counter() - t0)))
ss, _ = small.reset(pool[:512].to(dev))
ok = torch.equal(sb[:, :512], ss)
for s in range(T // 2):
    sb, rb = big.step(acts[s]); ss, rs = small.step(acts_pool[s][:, :512].to(dev))
    ok = ok and torch.equal(sb[:, :512], ss) and torch.equal(rb[:, :512], rs)
snap = big.snapshot()
big.rollout(acts[T // 2:], keep_all=True)          # warm-up: the first 3.5 GB output allocation is a cudaMalloc
big.restore(snap)
torch.cuda.synchronize(); t0 = time.perf_counter()
sb, rb = big.rollout(acts[T // 2:], keep_all=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
for s in range(T // 2, T):
    ss, rs = small.step(acts_pool[s][:, :512].to(dev))
ok = ok and torch.equal(sb[-1][:, :512], ss) and torch.equal(big.positions[:512], small.positions) and torch.equal(big.winners()[:512], small.winners())
print('chained rollout: %.1f us per step of %d env-steps -> %.3e env-steps/s; first 512 tracks identical to the small batch: %s' % (
    1e6 * dt / (T - T // 2), B * 4, (T - T // 2) * B * 4 / dt, ok))
print('memory allocated: %.1f GB' % (torch.cuda.max_memory_allocated() / 1e9))
