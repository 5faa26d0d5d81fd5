"""world_size-2 gloo tests of the sharding / gather plumbing (no GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from game_level_gan_b200 import dist as gdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        trials, boards, P = 3, 7, 2
        g = torch.Generator().manual_seed(0)
        tracks = torch.randn((trials * boards, 16, 2), generator=g)
        winners_all = torch.randint(-1, P, (trials * boards,), generator=g)       # what one big env would give
        local, (lo, hi) = gdist.shard_trial_major(tracks, trials, rank, world)
        assert local.shape == (trials * (hi - lo), 16, 2)
        assert torch.equal(local.view(trials, hi - lo, 16, 2), tracks.view(trials, boards, 16, 2)[:, lo:hi])
        w_local = winners_all.view(trials, boards)[:, lo:hi].reshape(-1)
        stats_local = torch.nn.functional.one_hot(w_local + 1, P + 1).view(trials, -1, P + 1).float().mean(0)
        stats = gdist.all_gather_winner_stats(stats_local, total=boards)
        ref = torch.nn.functional.one_hot(winners_all + 1, P + 1).view(trials, -1, P + 1).float().mean(0)
        assert torch.equal(stats, ref)
        # plain winners with unequal shards
        lo2, hi2 = gdist.shard_bounds(winners_all.numel(), rank, world)
        got = gdist.all_gather_winners(winners_all[lo2:hi2].clone(), total=winners_all.numel())
        assert torch.equal(got, winners_all)
        # the reusable gatherer: unequal shards (21 rows over 2 ranks), buffers allocated once, called twice
        sg = gdist.ShardGather(boards, (P + 1,), torch.float32, torch.device('cpu'))
        assert sg.sizes == [4, 3] and sg.local == hi - lo
        for rep in range(2):
            assert torch.equal(sg(stats_local + rep), ref + rep)
        eq_sg = gdist.ShardGather(8, (), torch.int8, torch.device('cpu'))     # equal shards: no index pass
        assert eq_sg.index is None
        assert torch.equal(eq_sg(torch.arange(4, dtype=torch.int64) + 4 * rank).long(), torch.arange(8))
        fin = (winners_all[lo2:hi2] >= 0)
        assert abs(gdist.finish_rate(fin) - float((winners_all >= 0).float().mean())) < 1e-6
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [gdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_gather_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: 'ok', 1: 'ok'}, results


def test_bind_host_to_gpu_is_harmless_without_a_gpu():
    """No NVML device here: nothing changes and None comes back; on a GPU box the set returned is a subset of the
    cores the process was allowed to use."""
    import os
    from game_level_gan_b200 import dist as gdist
    before = os.sched_getaffinity(0)
    got = gdist.bind_host_to_gpu(0)
    try:
        assert got is None or (got and got <= before)
    finally:
        os.sched_setaffinity(0, before)
