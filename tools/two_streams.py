"""Upper bound of overlapping consecutive launches: R replicas stepped on S streams round-robin."""
import sys, os, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
R = int(os.environ.get("R", 8))
if len(sys.argv) > 1: bench.B_TRACKS = int(sys.argv[1])
reps = [bench.Replica(i, 0, dev, 'fast') for i in range(R)]
for S in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(S)]
    def run(k):
        for i in range(k):
            with torch.cuda.stream(streams[i % S]):
                reps[i % R].cycle(100)
    run(R); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    K = 64
    run(K)
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('streams', S, 'us/launch', 1e3 * ms / (K * 100), 'env-steps/s %.3e' % (K * 100 * bench.B_TRACKS * 2 / (ms * 1e-3)))
