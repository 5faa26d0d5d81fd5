"""Timeline of one Race.host_rollout call (bench.py's e2e): when each chunk's kernel and D2H start / end on the GPU."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from game_level_gan_b200.games import Race, RaceConfig
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 25
env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=dev)
tape, snap, _ = bench.record_tape(env, bench.synthetic_tracks(4096, 5), 6, dev)
host = tape.cpu().pin_memory()[bench.PREROLL:bench.PREROLL + 100]
hr = env.host_rollout(100, chunk=chunk)
for _ in range(3):
    env.restore(snap); hr.run(host)
# instrumented copy of HostRollout.run
main = torch.cuda.current_stream(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
env.restore(snap)
torch.cuda.synchronize()
t0 = ev(); t0.record(main)
marks = []
hr.s_in.wait_stream(main); hr.s_out.wait_stream(main)
for c, (lo, hi) in enumerate(hr.bounds):
    with torch.cuda.stream(hr.s_in):
        hr.actions_d[lo:hi].copy_(hr.actions_h[lo:hi], non_blocking=True)
        hr.ev_in[c].record(hr.s_in)
for c, (lo, hi) in enumerate(hr.bounds):
    main.wait_event(hr.ev_in[c])
    k0 = ev(); k0.record(main)
    hr.plans[c].run()
    k1 = ev(); k1.record(main)
    hr.ev_done[c].record(main)
    with torch.cuda.stream(hr.s_out):
        hr.s_out.wait_event(hr.ev_done[c])
        d0 = ev(); d0.record(hr.s_out)
        hr.states_h[lo:hi].copy_(hr.states_d[lo:hi], non_blocking=True)
        hr.rewards_h[lo:hi].copy_(hr.rewards_d[lo:hi], non_blocking=True)
        d1 = ev(); d1.record(hr.s_out)
    marks.append((k0, k1, d0, d1, (hi - lo)))
main.wait_stream(hr.s_out)
t1 = ev(); t1.record(main)
torch.cuda.synchronize()
print('chunk %d: total %.3f ms' % (chunk, t0.elapsed_time(t1)))
for c, (k0, k1, d0, d1, n) in enumerate(marks):
    mb = n * 2 * 4096 * 21 * 4 / 1e6
    print('  chunk %d: kernel %.3f -> %.3f ms (%.3f), D2H %.3f -> %.3f ms (%.3f ms, %.1f GB/s)' % (
        c, t0.elapsed_time(k0), t0.elapsed_time(k1), k0.elapsed_time(k1), t0.elapsed_time(d0), t0.elapsed_time(d1),
        d0.elapsed_time(d1), mb / d0.elapsed_time(d1)))
