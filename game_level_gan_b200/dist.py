"""Multi-GPU plumbing of the Race path: tracks are sharded across ranks, one collective per episode.

Tracks are fully independent (no cross-track term in reset/step/winners), so every
rank steps its own contiguous block of boards with no data-path communication.  The only exchange is
the all-gather of the per-track winners / per-board winner statistics that the winner discriminator
consumes after an episode (train-gan.py:98, 103-105).  Works with the `nccl` backend on GPUs
(NVLink / NVSwitch) and with `gloo` on CPU tensors (tests).

Shard sizes follow from `shard_bounds` on every rank, so nothing about sizes is communicated and the
gather never synchronises with the host: one `all_gather_into_tensor` on buffers that a `ShardGather`
allocates once (shards padded to the largest, the padding dropped by a fixed index on the way out).
"""
import torch
import torch.distributed as dist


def shard_bounds(num_boards, rank, world):
    """Contiguous block [lo, hi) of boards owned by `rank`; blocks differ by at most one board."""
    base, extra = divmod(num_boards, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_trial_major(tracks, trials, rank, world):
    """train-gan.py:84 repeats the boards trial-major ([trials * boards, L, 2]).  Shard by BOARD so
    that all trials of a board live on one rank (the mean over trials at :104 is then local).
    Returns (local tracks, trial-major again, [trials * local_boards, L, 2]), (lo, hi))."""
    boards = tracks.size(0) // trials
    lo, hi = shard_bounds(boards, rank, world)
    local = tracks.view(trials, boards, *tracks.shape[1:])[:, lo:hi]
    return local.reshape(trials * (hi - lo), *tracks.shape[1:]).contiguous(), (lo, hi)


def _world(group):
    return dist.get_world_size(group) if dist.is_initialized() else 1


class ShardGather(object):
    """All-gather of per-board rows ([boards_local, *row] on every rank -> [boards, *row] in rank order) for a FIXED
    sharding of `total` boards by `shard_bounds`.  Buffers are allocated once; `__call__` is one collective and two
    device copies, no host synchronisation - safe inside a timed region or a CUDA graph."""

    def __init__(self, total, row_shape, dtype, device, group=None):
        self.group, self.total = group, int(total)
        self.world = _world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        spans = [shard_bounds(self.total, r, self.world) for r in range(self.world)]
        self.sizes = [hi - lo for lo, hi in spans]
        self.local = self.sizes[self.rank]
        self.cap = max(self.sizes) if self.sizes else 0
        row_shape = tuple(row_shape)
        self.wire = torch.zeros((self.cap,) + row_shape, dtype=dtype, device=device)
        self.recv = torch.empty((self.world * self.cap,) + row_shape, dtype=dtype, device=device)
        self.out = torch.empty((self.total,) + row_shape, dtype=dtype, device=device)
        if all(s == self.cap for s in self.sizes):
            self.index = None                      # equal shards: the receive buffer already is the result
        else:
            self.index = torch.cat([torch.arange(r * self.cap, r * self.cap + s) for r, s in enumerate(self.sizes)]
                                   ).to(device)

    def __call__(self, rows):
        if rows.size(0) != self.local:
            raise ValueError('this rank owns %d boards, got %d rows' % (self.local, rows.size(0)))
        if self.world == 1:
            self.out.copy_(rows)
            return self.out
        self.wire[:self.local].copy_(rows)                       # (casts to the wire dtype)
        dist.all_gather_into_tensor(self.recv, self.wire, group=self.group)
        if self.index is None:
            return self.recv
        torch.index_select(self.recv, 0, self.index, out=self.out)
        return self.out


def all_gather_winners(winners, total=None, group=None):
    """winners [b_local] int64 of every rank -> [total] int64 on every rank, in rank order.  On the wire the
    winners travel as int8 (values -1..P-1), 1 byte per track: 1 MB for 2^20 tracks, i.e. latency-bound on NVLink.
    `total` = number of tracks over all ranks, sharded by `shard_bounds` (default: world x b_local, equal shards).
    One-off convenience; loops should keep a `ShardGather`."""
    world = _world(group)
    if world == 1:
        return winners
    total = world * winners.numel() if total is None else total
    g = ShardGather(total, (), torch.int8, winners.device, group)
    return g(winners).to(torch.int64)


def all_gather_winner_stats(stats, total=None, group=None):
    """Per-board soft labels [boards_local, P+1] f32 of every rank -> [boards, P+1] in rank order
    (`total` boards sharded by `shard_bounds`; default: equal shards)."""
    world = _world(group)
    if world == 1:
        return stats
    total = world * stats.size(0) if total is None else total
    g = ShardGather(total, stats.shape[1:], stats.dtype, stats.device, group)
    return g(stats).clone()


def finish_rate(finishes, group=None):
    """Global mean of `game.finishes.float()` (train-gan.py:98) over all shards."""
    s = torch.stack((finishes.float().sum(), torch.tensor(float(finishes.numel()), device=finishes.device)))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    return (s[0] / s[1].clamp(min=1.)).item()


def bind_host_to_gpu(device_index):
    """One process per GPU: run this rank's host threads on the CPU cores NVML reports as local to its GPU (same NUMA
    node / PCIe root), so that the pinned staging buffers it allocates afterwards (first touch) and its copy threads
    sit next to the GPU - host<->device copies of eight ranks otherwise funnel through one socket.  Only cores the
    process is already allowed to use are kept; returns the new affinity set, or None when nothing was changed
    (NVML missing, no information, or no allowed core is local to the GPU)."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(device_index))
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    except Exception:  # noqa: BLE001
        return None
    local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
    allowed = os.sched_getaffinity(0)
    cores = local & allowed
    if not cores or cores == allowed:
        return None
    os.sched_setaffinity(0, cores)
    return cores
