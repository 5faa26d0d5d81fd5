"""GPU time of a LONE step launch: graphs of 1 and 9 strictly serialised step kernels (no PDL inside a captured graph)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceConfig
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
tape, snap, _ = bench.record_tape(env, bench.synthetic_tracks(4096, 5), 6, dev)
env.restore(snap)
P, B = 2, 4096
states = torch.zeros((P, B, 20), device=dev); rewards = torch.zeros((P, B), device=dev)
base = torch.tensor([env.steps, env._seq, 2 ** 31 - 1], dtype=torch.int32, device=dev)
acts = tape[100:110].contiguous()
def make(n):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        for j in range(n):
            env.step_into(acts[j], states, rewards, base, j + 1)
    return g
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for variant in ('fast', 'warp'):
    env.variant = variant
    for n in (1, 9):
        g = make(n)
        tot, cnt = 0., 0
        for r in range(60):
            env.restore(snap); torch.cuda.synchronize()
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            if r >= 10: tot += e0.elapsed_time(e1); cnt += 1
        res[(variant, n)] = 1e3 * tot / cnt
    print('%s: graph of 1 kernel %.2f us, of 9 kernels %.2f us -> %.2f us per serialised launch' % (
        variant, res[(variant, 1)], res[(variant, 9)], (res[(variant, 9)] - res[(variant, 1)]) / 8))
