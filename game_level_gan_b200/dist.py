"""Multi-GPU plumbing of the Race path: tracks are sharded across ranks, one collective per episode.

Tracks are fully independent (no cross-track term in reset/step/winners), so every
rank steps its own contiguous block of boards with no data-path communication.  The only exchange is
the all-gather of the per-track winners (and, optionally, the finish counts) that the winner
discriminator consumes after an episode (train-gan.py:98, 103-105).  Works with the `nccl` backend
on GPUs (NVLink / NVSwitch) and with `gloo` on CPU tensors (tests).
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


def all_gather_winners(winners, group=None):
    """winners [b_local] int64 of every rank -> [sum b_local] on every rank, in rank order.

    On the wire the winners travel as int8 (values -1..P-1), 1 byte per track: 1 MB for 2^20 tracks,
    i.e. latency-bound on NVLink; shards may have different sizes (padded to the largest)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return winners
    world = dist.get_world_size(group)
    n = torch.tensor([winners.numel()], dtype=torch.int64, device=winners.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    wire = torch.full((cap,), -1, dtype=torch.int8, device=winners.device)
    wire[:winners.numel()] = winners.to(torch.int8)
    out = torch.empty((world * cap,), dtype=torch.int8, device=winners.device)
    dist.all_gather_into_tensor(out, wire, group=group)
    parts = [out[r * cap:r * cap + sizes[r]] for r in range(world)]
    return torch.cat(parts).to(torch.int64)


def all_gather_winner_stats(stats, group=None):
    """Per-board soft labels [boards_local, P+1] f32 of every rank -> [boards, P+1] in rank order
    (equal shard sizes are not required)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    n = torch.tensor([stats.size(0)], dtype=torch.int64, device=stats.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes)
    wire = torch.zeros((cap, stats.size(1)), dtype=stats.dtype, device=stats.device)
    wire[:stats.size(0)] = stats
    out = torch.empty((world * cap, stats.size(1)), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, wire, group=group)
    return torch.cat([out[r * cap:r * cap + sizes[r]] for r in range(world)])


def finish_rate(finishes, group=None):
    """Global mean of `game.finishes.float()` (train-gan.py:98) over all shards."""
    s = torch.stack((finishes.float().sum(), torch.tensor(float(finishes.numel()), device=finishes.device)))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    return (s[0] / s[1].clamp(min=1.)).item()
