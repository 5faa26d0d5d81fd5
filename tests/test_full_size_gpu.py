"""Parity at the sizes BASELINE.json names: config 2 directly against the C oracle,
config 4 (one GPU's shard: 131 072 tracks x 4 cars) and config 5 (Pacman, 65 536 boards) through
size-independent properties - tracks / boards do not interact, so a large batch assembled from a small
pool must reproduce, car by car and bit for bit, the results of the pool, which itself is checked
against the oracle."""
import ctypes

import numpy as np
import pytest
import torch

from tests.helpers import eq, nmismatch

pytestmark = pytest.mark.gpu

CARS4 = [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)]


def iid9_tracks(n, gen):
    t = torch.zeros(n, 128, 2)
    t[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (n, 128), generator=gen)]
    return t


def biased_actions(T, P, B, gen, p_forward=0.5):
    a = torch.randint(0, 9, (T, P, B), generator=gen)
    return torch.where(torch.rand((T, P, B), generator=gen) < p_forward, torch.ones_like(a), a)


def c_oracle(cars, framerate=1. / 20., timeout=40.):
    from game_level_gan_b200.games import _tables
    from oracle import c_oracle as co
    from oracle import race_oracle as ro
    pr = _tables.race_params([ro.Car(*c) for c in cars], framerate, timeout, 18, 10.)
    out = co.RaceParams()
    ctypes.memmove(ctypes.byref(out), ctypes.byref(pr), ctypes.sizeof(out))
    return co.CRace(out)


def test_config2_full_size_vs_c_oracle():
    """4096 generator-like tracks x 2 cars, 40 steps: every output of every step against the C oracle."""
    from game_level_gan_b200.games import Race, RaceCar, _tables
    g = torch.Generator().manual_seed(2024)
    B, P, T = 4096, 2, 40
    tracks, acts = iid9_tracks(B, g), biased_actions(T, P, B, g)
    orc = c_oracle(CARS4[:P])
    st, ct, _ = _tables.heading_tables(128)
    so, _ = orc.reset(tracks.numpy(), st.numpy(), ct.numpy())
    env = Race(timeout=40., cars=[RaceCar(*c) for c in CARS4[:P]], framerate=1. / 20., log_history=False)
    s0, _ = env.reset(tracks)
    bad = nmismatch(s0, so) + nmismatch(env._valid_tracks, orc.valid)
    for s in range(T):
        so, ro_ = orc.step(acts[s].numpy())
        sg, rg = env.step(acts[s].cuda())
        bad += nmismatch(sg, so) + nmismatch(rg, ro_) + nmismatch(env.positions, orc.pos)
        bad += nmismatch(env._alive, orc.alive) + nmismatch(env.scores, orc.scores)
    assert bad == 0
    assert eq(env.winners(), orc.winners())


def test_config2_full_size_100_step_fused_rollout_vs_c_oracle():
    """The bench's production path at the bench's size: ONE fused launch plays 100 steps of 4096 tracks x 2 cars
    (`north_star`: 100-step rollouts must be checked); every observation and reward of every step and the final car
    state against the C oracle, bit for bit.  A second rollout continues the episode through the time when cars die and
    finish."""
    from game_level_gan_b200.games import Race, RaceCar, _tables
    g = torch.Generator().manual_seed(4048)
    B, P, T = 4096, 2, 100
    tracks = iid9_tracks(B, g)
    # an action tape on which most cars survive (uniformly random actions kill every car within ~100 steps, and then the
    # reference returns 19-wide zeros): a wall-avoiding driver plays the episode once, closed loop, 10 % random actions
    drv = Race(timeout=40., cars=[RaceCar(*c) for c in CARS4[:P]], framerate=1. / 20., log_history=False)
    obs, _ = drv.reset(tracks)
    gd = torch.Generator(device='cuda').manual_seed(7)
    tape = []
    for _ in range(2 * T):
        left, right = obs[..., 7] + obs[..., 8] + 0.5 * obs[..., 6], obs[..., 10] + obs[..., 11] + 0.5 * obs[..., 12]
        steer = torch.zeros(obs.shape[:2], dtype=torch.int64, device='cuda')
        steer[right > left + 0.02] = 1
        steer[left > right + 0.02] = 2
        thr = torch.where(obs[..., 18] > 0.05, 0, 1)
        a = steer * 3 + thr
        rnd = torch.rand(a.shape, generator=gd, device='cuda') < 0.1
        a = torch.where(rnd, torch.randint(0, 9, a.shape, generator=gd, device='cuda'), a)
        tape.append(a)
        obs, _ = drv.step(a)
    acts = torch.stack(tape).cpu()
    orc = c_oracle(CARS4[:P])
    st, ct, _ = _tables.heading_tables(128)
    so, _ = orc.reset(tracks.numpy(), st.numpy(), ct.numpy())
    env = Race(timeout=40., cars=[RaceCar(*c) for c in CARS4[:P]], framerate=1. / 20., log_history=False)
    s0, _ = env.reset(tracks)
    assert nmismatch(s0, so) == 0
    for half in range(2):
        a = acts[half * T:(half + 1) * T]
        plan = env.rollout_plan(a.cuda(), keep_all=True, mode='fused')
        assert plan.launches == 1
        sg, rg = plan.run()
        sg, rg = sg.cpu(), rg.cpu()
        bad = 0
        for s in range(T):
            so, ro_ = orc.step(a[s].numpy())
            bad += nmismatch(sg[s], so) + nmismatch(rg[s], ro_)
        bad += nmismatch(env.positions, orc.pos) + nmismatch(env.directions, orc.dir) + nmismatch(env.speeds, orc.speed)
        bad += nmismatch(env._alive, orc.alive) + nmismatch(env._finishes, orc.finishes) + nmismatch(env.scores, orc.scores)
        assert bad == 0, 'rollout %d' % half
    assert eq(env.winners(), orc.winners())
    assert 0 < int(env._alive.sum()) < B * P and int(env._alive.sum()) > B     # most cars drove the 200 steps, some died


def test_config4_shard_is_a_gather_of_a_small_batch():
    """131 072 tracks x 4 cars (config 4 on 8 GPUs, per GPU) built by repeating 256 distinct tracks in a
    random order: all 524 288 cars must equal, bit for bit, the corresponding cars of the 256-track batch
    (which is compared with the C oracle here), step by step and through a chained rollout."""
    from game_level_gan_b200.games import Race, RaceCar, _tables
    g = torch.Generator().manual_seed(4)
    pool, B, P, T = 256, 131072, 4, 24
    base, base_acts = iid9_tracks(pool, g), biased_actions(T, P, pool, g, 0.6)
    idx = torch.randint(0, pool, (B,), generator=g)
    idx[:pool] = torch.arange(pool)
    cars = [RaceCar(*c) for c in CARS4]
    small = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
    big = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
    orc = c_oracle(CARS4)
    st, ct, _ = _tables.heading_tables(128)
    so, _ = orc.reset(base.numpy(), st.numpy(), ct.numpy())
    ss, _ = small.reset(base)
    sb, any_valid = big.reset(base[idx])
    assert any_valid and eq(ss, so) and eq(sb, ss[:, idx.cuda()])
    assert eq(big._valid_tracks, small._valid_tracks[idx.cuda()])
    dev_idx = idx.cuda()
    half = T // 2
    for s in range(half):
        so, ro_ = orc.step(base_acts[s].numpy())
        ss, rs = small.step(base_acts[s].cuda())
        sb, rb = big.step(base_acts[s][:, idx].cuda())
        assert eq(ss, so) and eq(rs, ro_), 'small batch vs oracle, step %d' % s
        assert eq(sb, ss[:, dev_idx]) and eq(rb, rs[:, dev_idx]), 'big batch, step %d' % s
    # second half as one chained rollout on the big batch
    for s in range(half, T):
        ss, rs = small.step(base_acts[s].cuda())
    sb, rb = big.rollout(base_acts[half:][:, :, idx].cuda())
    assert eq(sb, ss[:, dev_idx]) and eq(rb, rs[:, dev_idx])
    for a, b in ((big.positions, small.positions), (big.directions, small.directions), (big.speeds, small.speeds),
                 (big._alive, small._alive), (big._finishes, small._finishes), (big.scores, small.scores)):
        assert eq(a, b[dev_idx])
    assert eq(big.winners(), small.winners()[dev_idx])
    assert big.finished() == small.finished()


def test_config5_pacman_full_size_replication():
    """65 536 boards of 15x15 with 2 players built from 128 distinct boards: every board evolves exactly like
    its source board in the 128-board batch, which is compared with the numpy oracle."""
    from game_level_gan_b200.games import Pacman
    from oracle.pacman_oracle import PacmanOracle
    rng = np.random.default_rng(5)
    pool, B, H, W, P, T = 128, 65536, 15, 15, 2, 12
    fields = rng.choice(4, size=(pool, H, W), p=[0.4, 0.5, 0.07, 0.03])
    board = np.zeros((pool, H, W, 4 + P), dtype=np.int32)
    np.put_along_axis(board[..., :4], fields[..., None], 1, axis=-1)
    for p, (x, y) in enumerate(((0, 0), (H - 1, W - 1))):
        board[:, x, y, :] = 0
        board[:, x, y, 0] = 1
        board[:, x, y, 4 + p] = 1
    idx = rng.integers(0, pool, size=B)
    idx[:pool] = np.arange(pool)
    orc = PacmanOracle((H, W), P, pool)
    orc.reset(board.copy())
    small = Pacman((H, W), P, batch_size=pool)
    big = Pacman((H, W), P, batch_size=B)
    small.reset(board.copy())
    big.reset(board[idx].copy())
    for t in range(T):
        acts = rng.integers(0, 5, size=(pool, P)).astype(np.int32)
        _, orew = orc.step(acts)
        _, srew = small.step(acts)
        _, brew = big.step(acts[idx])
        assert np.array_equal(np.stack(srew), np.stack(orew))
        assert np.array_equal(np.stack(brew), np.stack(srew)[:, idx])
    assert np.array_equal(small.grid, orc.grid)
    assert np.array_equal(big.grid, small.grid[idx])
