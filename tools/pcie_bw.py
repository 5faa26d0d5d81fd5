"""Pinned-memory copy bandwidth of the box (what bounds bench.py's e2e): D2H / H2D of several sizes, one and two streams."""
import torch, time
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
def bw(nbytes, direction, streams=1, reps=10):
    hs = [torch.empty(nbytes // streams, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes // streams, dtype=torch.uint8, device=dev) for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go():
        for h, d, s in zip(hs, ds, ss):
            with torch.cuda.stream(s):
                (h if direction == 'd2h' else d).copy_(d if direction == 'd2h' else h, non_blocking=True)
    go(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): go()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
for direction in ('d2h', 'h2d'):
    for n in (1 << 20, 16 << 20, 64 << 20, 256 << 20):
        print('%s %4d MB: 1 stream %.1f GB/s, 2 streams %.1f GB/s' % (direction, n >> 20, bw(n, direction, 1), bw(n, direction, 2)))
# both directions at once
h1 = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d1 = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
h2 = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('bidirectional 64 MB each way: %.1f GB/s per direction' % ((64 << 20) * 10 / dt / 1e9))
