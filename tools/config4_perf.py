"""Config-4-shaped throughput on one GPU: B tracks x 4 cars, chained rollouts, all cars alive."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceCar
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
R = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)]]
bench.B_TRACKS = B
reps = []
for i in range(R):
    env = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, device=dev)
    acts, snap, alive_end = bench.record_tape(env, bench.synthetic_tracks(B, 100 + i), 200 + i, dev)
    reps.append((env, acts, snap, alive_end))
def cycle(i):
    env, acts, snap, _ = reps[i % R]
    env.restore(snap); env.rollout(acts[100:200], keep_all=True)
for i in range(R): cycle(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 4 * R
e0.record()
for i in range(K): cycle(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
steps = K * 100
print('B=%d P=4: %.2f us/launch, %.3e env-steps/s, alive at tape end %.3f; roofline 925 B/env-step -> frac %.3f' % (
    B, 1e3 * ms / steps, steps * B * 4 / (ms * 1e-3), sum(r[3] for r in reps) / R, steps * B * 4 * 925 / (ms * 1e-3) / 6549.8e9))
