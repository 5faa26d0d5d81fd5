"""Reset cost (geometry build, validity, extents, init, first step) at config 2 and at a config-4 shard."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceConfig
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
for B in (4096, 131072):
    tracks = bench.synthetic_tracks(B, 3).to(dev)
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        env.reset(tracks)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('B=%d: reset %.2f ms (%.1f ns per track) valid fraction %.3f' % (B, 1e3 * dt, 1e9 * dt / B, float(env._valid_tracks.float().mean())))

# the two validity kernels alone (sign matrix vs the loop over all pairs)
from game_level_gan_b200 import _lib
from game_level_gan_b200._lib import ptr
for B in (4096, 131072):
    tracks = bench.synthetic_tracks(B, 3).to(dev)
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
    env.reset(tracks)
    out = torch.empty(B, dtype=torch.uint8, device=dev)
    for name in ('glg_track_validate', 'glg_track_validate_pairs'):
        fn = getattr(_lib.lib(), name)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(4):
            if r == 1: e0.record()
            fn(ptr(env._geom), B, env._geom.size(2), ptr(out), _lib.stream_ptr(dev))
        e1.record(); torch.cuda.synchronize()
        print('B=%d %s: %.1f us per call, equal to reset: %s' % (B, name, 1e3 * e0.elapsed_time(e1) / 3, bool((out == env._valid_tracks).all())))
