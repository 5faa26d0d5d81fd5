"""Parity of the CUDA Race path (through the C ABI) against the reference-generated fixtures, the
torch oracle and the C oracle.  These are the parity tests proper: run with `-m gpu` on a B200."""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests.helpers import RACE_CASES, eq, load_case, nmismatch, t

pytestmark = pytest.mark.gpu

QUANTISED = ['predef', 'iid9', 'loops', 'p1_crash', 'p4_short', 'agents']


def make_env(case, variant, log_history=False):
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in case['cars'].tolist()]
    return Race(timeout=float(case['timeout']), cars=cars, framerate=float(case['framerate']),
                log_history=log_history, variant=variant)


def replay(case, env, geometry=None, exact=True):
    """Teacher-forced replay of a fixture; returns the number of mismatching elements per field."""
    states, any_valid = env.reset(t(case['tracks']), geometry=geometry)
    bad = dict(states=0, rewards=0, pos=0, dir=0, speed=0, alive=0, finishes=0, scores=0, finished=0, width=0)
    assert any_valid == bool(case['any_valid'])
    bad['states'] += nmismatch(states, case['states'][0])
    bad['finished'] += int(env.finished() != bool(case['finished'][0]))
    for s in range(case['actions'].shape[0]):
        states, rewards = env.step(t(case['actions'][s]).cuda())
        w = int(case['widths'][s + 1])
        if states.size(-1) != w:
            bad['width'] += 1
            continue
        bad['states'] += nmismatch(states, case['states'][s + 1][:, :, :w])
        bad['rewards'] += nmismatch(rewards, case['rewards'][s])
        for k, v in (('pos', env.positions), ('dir', env.directions), ('speed', env.speeds),
                     ('alive', env.alive), ('finishes', env.finishes), ('scores', env.scores)):
            bad[k] += nmismatch(v, case[k][s + 1])
        bad['finished'] += int(env.finished() != bool(case['finished'][s + 1]))
    return bad


@pytest.mark.parametrize('variant', ['fast', 'warp', 'scan', 'brute'])
@pytest.mark.parametrize('name', QUANTISED)
def test_fixture_bit_exact_end_to_end(name, variant):
    """Generator-like tracks (arcs in 1/4 steps): geometry, every step output and the winners are
    bit-identical to the reference, building the geometry with our own kernel."""
    c = load_case(name)
    env = make_env(c, variant)
    bad = replay(c, env)
    assert eq(env.segments, c['centre']) and eq(env.left_vecs, c['left']) and eq(env.right_vecs, c['right'])
    assert eq(env.valid, c['valid'])
    assert all(v == 0 for v in bad.values()), bad
    assert eq(env.winners(), c['winners'])


@pytest.mark.parametrize('variant', ['fast', 'warp', 'scan', 'brute'])
def test_fixture_float_tracks(variant):
    """Arbitrary float arcs/widths: step parity is bit-exact on the reference's geometry; the build
    itself differs from the reference's SLEEF sin/cos by a few ulp (tolerance 2e-5 absolute)."""
    c = load_case('floatw')
    env = make_env(c, variant)
    bad = replay(c, env, geometry=(t(c['centre']), t(c['left']), t(c['right'])))
    assert all(v == 0 for v in bad.values()), bad
    assert eq(env.winners(), c['winners'])
    env2 = make_env(c, variant)
    env2.reset(t(c['tracks']))
    for k, v in (('centre', env2.segments), ('left', env2.left_vecs), ('right', env2.right_vecs)):
        np.testing.assert_allclose(v.cpu().numpy(), c[k], rtol=0, atol=2e-5)
    assert eq(env2.valid, c['valid'])


def test_attribute_layouts_and_history():
    """Public attributes in the reference's layouts (games/race.py:161-190) and the history list."""
    c = load_case('predef')
    env = make_env(c, 'fast', log_history=True)
    env.reset(t(c['tracks']))
    B, P, N = 12, 2, 130
    assert env.bounds.shape == (B * P, 2 * 129 + 1, 4) and env.reward_bound.shape == (B * P, 1, 4)
    assert env.line_bounds.shape == (B * P, 2 * N, 2) and env.line_bounds.device.type == 'cpu'
    assert env.left_bounds.shape == (B, 129, 4) and env.right_bounds.shape == (B, 129, 4)
    assert env.positions.shape == (B, P, 2) and env.alive.dtype == torch.bool and env.scores.dtype == torch.int32
    assert env.valid.shape == (B * P,) and env.valid.dtype == torch.bool
    assert env.state_shape() == (20,) and env.actions == 9 and env.num_players == 2
    from oracle import race_oracle as ro
    walls, finish = ro.wall_table(t(c['left']), t(c['right']))
    assert eq(env.bounds.view(B, P, -1, 4)[:, 1], walls) and eq(env.reward_bound.view(B, P, 1, 4)[:, 0], finish)
    for s in range(5):
        env.step(t(c['actions'][s]).cuda())
    h = env.history
    assert len(h) == 6                                   # reset's noop step + 5
    pos, dirs, acts, alive = h[-1]
    assert eq(torch.tensor(pos), c['pos'][5][0]) and eq(torch.tensor(dirs), c['dir'][5][0])
    assert alive == c['alive'][5][0].tolist()
    assert list(env.iterate_valid(['a'] * B)) == [(i, 'a') for i in range(B)]


def _c_oracle_for(env, case_cars, framerate, timeout):
    from oracle import c_oracle
    from game_level_gan_b200.games import _tables
    from oracle import race_oracle as ro
    pr = _tables.race_params([ro.Car(*c) for c in case_cars], framerate, timeout, 18, 10.)
    out = c_oracle.RaceParams()
    ctypes.memmove(ctypes.byref(out), ctypes.byref(pr), ctypes.sizeof(out))
    return c_oracle.CRace(out)


@pytest.mark.parametrize('P', [2, 4, 7])
def test_free_running_rollout_vs_c_oracle(P):
    """100-step free-running rollout on 512 iid-9 tracks, random + forward-biased actions: CUDA (fast and
    brute) against the C oracle, every step, every output - bit-exact."""
    from game_level_gan_b200.games import Race, RaceCar, _tables
    cars = [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.), (70., 2., 30.), (40., 4., 90.),
            (90., 1., 45.)][:P]
    g = torch.Generator().manual_seed(99)
    B, T = 512, 100
    space = torch.linspace(-1., 1., 9)
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = space[torch.randint(0, 9, (B, 128), generator=g)]
    acts = torch.randint(0, 9, (T, P, B), generator=g)
    acts = torch.where(torch.rand((T, P, B), generator=g) < 0.5, torch.ones_like(acts), acts)
    envs = {v: Race(timeout=40., cars=[RaceCar(*c) for c in cars], framerate=1. / 20., log_history=False,
                    variant=v) for v in ('fast', 'warp', 'scan', 'brute')}
    orc = _c_oracle_for(None, cars, 1. / 20., 40.)
    st, ct, _ = _tables.heading_tables(128)
    so, _ = orc.reset(tracks.numpy(), st.numpy(), ct.numpy())
    bad = {v: 0 for v in envs}
    for v, env in envs.items():
        s0, _ = env.reset(tracks)
        assert eq(env.right_vecs, orc.geom[:, 0]) and eq(env.left_vecs, orc.geom[:, 1]) and eq(env.segments, orc.geom[:, 2]), 'geometry'
        assert eq(env._valid_tracks, orc.valid)
        bad[v] += nmismatch(s0, so)
    for s in range(T):
        so, ro_ = orc.step(acts[s].numpy())
        for v, env in envs.items():
            sg, rg = env.step(acts[s].cuda())
            bad[v] += nmismatch(sg, so) + nmismatch(rg, ro_)
            bad[v] += nmismatch(env.positions, orc.pos) + nmismatch(env.speeds, orc.speed)
            bad[v] += nmismatch(env._alive, orc.alive) + nmismatch(env.scores, orc.scores)
    assert bad == {'fast': 0, 'warp': 0, 'scan': 0, 'brute': 0}, bad
    for env in envs.values():
        assert eq(env.winners(), orc.winners())
    assert int(orc.alive.sum()) < B * P            # the rollout really killed cars


def test_rollout_equals_steps_and_winner_stats():
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(5)
    trials, boards, T = 3, 40, 60
    space = torch.linspace(-1., 1., 9)
    base = torch.zeros(boards, 128, 2)
    base[:, :, 0] = space[torch.randint(0, 9, (boards, 128), generator=g)]
    tracks = base.repeat(trials, 1, 1)                  # trial-major, train-gan.py:84
    B = trials * boards
    acts = torch.randint(0, 9, (T, 2, B), generator=g)
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    a.reset(tracks)
    per_step = [a.step(acts[s].cuda()) for s in range(T)]
    for mode in ('fused', 'chained', 'stepwise'):
        b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        b.reset(tracks)
        states, rewards = b.rollout(acts.cuda(), keep_all=True, mode=mode)
        assert eq(states, torch.stack([s for s, _ in per_step])) and eq(rewards, torch.stack([r for _, r in per_step]))
        assert eq(a.positions, b.positions) and eq(a.scores, b.scores) and a.steps == b.steps
        assert a.finished() == b.finished()
        # rewind and replay in two chained pieces, keeping only the last outputs (same buffer every step)
        c = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        c.reset(tracks)
        snap = c.snapshot()
        c.rollout(acts[:17].cuda(), mode=mode)
        c.restore(snap)
        c.rollout(acts[:30].cuda(), mode=mode)
        s_last, r_last = c.rollout(acts[30:].cuda(), mode=mode)
        assert eq(s_last, per_step[-1][0]) and eq(r_last, per_step[-1][1]) and eq(c.positions, a.positions)
        assert c.finished() == a.finished()
    w = a.winners()
    ref = torch.nn.functional.one_hot(w.cpu() + 1, 3).view(trials, -1, 3).float().mean(0)
    assert eq(a.winner_stats(trials), ref)


def test_empty_batch_and_errors():
    from game_level_gan_b200.games import Race, RaceConfig
    from game_level_gan_b200._lib import GlgError
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20.)
    states, any_valid = env.reset(torch.zeros(0, 128, 2))
    assert states.shape == (2, 0, 19) and any_valid is False and env.finished()
    assert env.winners().shape == (0,)
    env.reset(torch.zeros(3, 128, 2))
    with pytest.raises(ValueError):
        env.step(torch.zeros(2, 4, dtype=torch.int64))
    with pytest.raises(GlgError):
        env.reset(torch.zeros(2, 600, 2))               # L beyond the kernel limit
    with pytest.raises(GlgError):
        Race(timeout=40., cars=RaceConfig.cars, device='cpu')


def test_reset_from_generator_levels():
    """geometry straight from the 4-bit arc levels of the discrete generator equals the geometry
    built from the float tracks (and therefore the reference's, fixture iid9), odd L included."""
    from game_level_gan_b200.games import Race, RaceConfig
    c = load_case('iid9')
    tracks = t(c['tracks'])
    levels = torch.round(tracks[:, :, 0] * 4).long() + 4
    assert int(levels.min()) >= 0 and int(levels.max()) <= 8
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    states, any_valid = env.reset_levels(levels)
    assert eq(env.segments, c['centre']) and eq(env.left_vecs, c['left']) and eq(env.right_vecs, c['right'])
    assert eq(env.valid, c['valid']) and eq(states, c['states'][0]) and any_valid == bool(c['any_valid'])
    g = torch.Generator().manual_seed(3)
    lv = torch.randint(0, 9, (33, 77), generator=g)
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    tr = torch.zeros(33, 77, 2)
    tr[:, :, 0] = torch.linspace(-1., 1., 9)[lv]
    sa, _ = a.reset(tr)
    sb, _ = b.reset_levels(lv)
    assert eq(a._geom, b._geom) and eq(sa, sb) and eq(a._valid_tracks, b._valid_tracks)
    # odd L: the record is not 16-byte granular, the step kernels take their plain-load path
    c2 = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant='brute')
    sc, _ = c2.reset(tr)
    assert eq(sa, sc)
    for s in range(25):
        acts = torch.randint(0, 9, (2, 33), generator=g)
        acts[:, ::2] = 1
        (sa, ra), (sc, rc) = a.step(acts.cuda()), c2.step(acts.cuda())
        assert eq(sa, sc) and eq(ra, rc) and eq(a.positions, c2.positions) and eq(a.scores, c2.scores)


@pytest.mark.parametrize('P', [1, 3, 4, 8])
def test_chained_rollout_other_player_counts(P):
    """Fused (one persistent kernel) and chained rollouts (LL hand-over with keep_all, release/acquire stamps
    without) for 1, 3, 4 and 8 cars per track - an idle half warp, several warps per track - against per-step calls,
    which are themselves compared with the C oracle."""
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.), (70., 2., 30.),
                                  (40., 4., 90.), (90., 1., 45.), (55., 2., 70.)][:P]]
    g = torch.Generator().manual_seed(40 + P)
    B, T = 301, 45
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    acts = torch.randint(0, 9, (T, P, B), generator=g)
    acts = torch.where(torch.rand((T, P, B), generator=g) < 0.5, torch.ones_like(acts), acts)
    a = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
    s0, _ = a.reset(tracks)
    per_step = [a.step(acts[s].cuda()) for s in range(T)]
    # the per-step results themselves against the C oracle (pinned to the reference's fixtures, tests/test_oracle_c.py):
    # the rollout modes below are then held to reference-equivalent values, not only to another CUDA path
    import ctypes
    from game_level_gan_b200.games import _tables
    from oracle import c_oracle as co
    cpr = co.RaceParams()
    ctypes.memmove(ctypes.byref(cpr), ctypes.byref(_tables.race_params(cars, 1. / 20., 40., 18, 10.)), ctypes.sizeof(cpr))
    orc = co.CRace(cpr)
    st_, ct_, _ = _tables.heading_tables(128)
    so, _ = orc.reset(tracks.numpy(), st_.numpy(), ct_.numpy())
    assert eq(s0, so)
    for s in range(T):
        so, ro_ = orc.step(acts[s].numpy())
        if so.shape[-1] != 20:                       # nobody alive any more: the reference's 19-wide early-out
            break
        assert eq(per_step[s][0], so) and eq(per_step[s][1], ro_), ('oracle', s)
    assert s >= 20
    for keep_all, mode in ((True, 'fused'), (False, 'fused'), (True, 'chained'), (False, 'chained'), (True, 'stepwise')):
        b = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
        b.reset(tracks)
        states, rewards = b.rollout(acts.cuda(), keep_all=keep_all, mode=mode)
        if keep_all:
            assert eq(states, torch.stack([s for s, _ in per_step])) and eq(rewards, torch.stack([r for _, r in per_step]))
        else:
            assert eq(states, per_step[-1][0]) and eq(rewards, per_step[-1][1])
        for x, y in ((a.positions, b.positions), (a.directions, b.directions), (a.speeds, b.speeds),
                     (a.alive, b.alive), (a.finishes, b.finishes), (a.scores, b.scores)):
            assert eq(x, y)
        assert a.finished() == b.finished() and eq(a.winners(), b.winners())


@pytest.mark.parametrize('O,max_distance', [(8, 10.), (9, 6.), (32, 10.), (18, 3.)])
def test_other_ray_counts_and_ranges(O, max_distance):
    """observation_size / max_distance other than the defaults (games/race.py:26): even ray counts run the
    single-pass pruning kernel, odd ones the literal loop; both against the torch oracle and the C oracle."""
    from game_level_gan_b200.games import Race, RaceCar, _tables
    from oracle import c_oracle
    from oracle import race_oracle as ro
    cars = [(60., 4., 40.), (60., 1., 80.)]
    g = torch.Generator().manual_seed(1000 + O)
    B, T = 96, 40
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    acts = torch.randint(0, 9, (T, 2, B), generator=g)
    acts = torch.where(torch.rand((T, 2, B), generator=g) < 0.5, torch.ones_like(acts), acts)
    pr = _tables.race_params([ro.Car(*c) for c in cars], 1. / 20., 40., O, max_distance)
    cpr = c_oracle.RaceParams()
    ctypes.memmove(ctypes.byref(cpr), ctypes.byref(pr), ctypes.sizeof(cpr))
    orc = c_oracle.CRace(cpr)
    st, ct, _ = _tables.heading_tables(128)
    so, _ = orc.reset(tracks.numpy(), st.numpy(), ct.numpy())
    tor = ro.RaceOracle(timeout=40., cars=[ro.Car(*c) for c in cars], observation_size=O, max_distance=max_distance,
                        framerate=1. / 20.)
    s_t, _ = tor.reset(tracks)
    envs = {v: Race(timeout=40., cars=[RaceCar(*c) for c in cars], observation_size=O, max_distance=max_distance,
                    framerate=1. / 20., log_history=False, variant=v) for v in ('fast', 'brute')}
    for v, env in envs.items():
        s0, _ = env.reset(tracks)
        assert s0.shape == (2, B, O + 2) and eq(s0, so) and eq(s0, s_t), v
    for s in range(T):
        so, ro_ = orc.step(acts[s].numpy())
        s_t, r_t = tor.step(acts[s]) if s < 12 else (None, None)          # the torch oracle is slow: first steps only
        for v, env in envs.items():
            sg, rg = env.step(acts[s].cuda())
            assert eq(sg, so) and eq(rg, ro_), (v, s)
            if s_t is not None:
                assert eq(sg, s_t) and eq(rg, r_t), (v, s)
            assert eq(env.positions, orc.pos) and eq(env.scores, orc.scores)


@pytest.mark.parametrize('P', [2, 3])
def test_unpruned_fallback_paths(P):
    """Cars for which a precondition of the pruning fails (heading norm far from 1, car far from the origin) take
    the brute-force path INSIDE the production kernels: force that state and compare packed / warp / scan with
    the literal kernel, step by step and through a chained rollout."""
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.)][:P]]
    g = torch.Generator().manual_seed(77 + P)
    B, T = 64, 24
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    acts = torch.randint(0, 9, (T, P, B), generator=g)
    acts = torch.where(torch.rand((T, P, B), generator=g) < 0.5, torch.ones_like(acts), acts)
    envs = {v: Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, variant=v)
            for v in ('brute', 'fast', 'warp', 'scan')}
    for env in envs.values():
        env.reset(tracks)
        for s in range(4):
            env.step(acts[s].cuda())
        # headings with |d|^2 = 2.56 (outside (0.5, 2)) on every 3rd track, cars moved 300 units away on every 5th
        env.directions[::3] *= 1.6
        env.positions[::5] += 300.
        env.directions[1::7] *= 0.5
    ref = envs['brute']
    for s in range(4, 14):
        so, ro_ = ref.step(acts[s].cuda())
        for v in ('fast', 'warp', 'scan'):
            sg, rg = envs[v].step(acts[s].cuda())
            assert eq(sg, so) and eq(rg, ro_), (v, s)
            assert eq(envs[v].positions, ref.positions) and eq(envs[v].alive, ref.alive) and eq(envs[v].scores, ref.scores)
    so, ro_ = ref.rollout(acts[14:].cuda(), keep_all=True)
    for v in ('fast', 'warp', 'scan'):
        sg, rg = envs[v].rollout(acts[14:].cuda(), keep_all=True)
        assert eq(sg, so) and eq(rg, ro_), v
        assert eq(envs[v].positions, ref.positions) and eq(envs[v].winners(), ref.winners())
    assert float(so[..., :18].abs().sum()) > 0.


def test_degenerate_car_on_the_wall_line():
    """Straight tracks with cars placed exactly ON a boundary line (and on the start line): every collinear wall
    becomes a candidate for all rays (queue overflow -> in-kernel brute force), the forward / backward rays are
    parallel to the walls they touch (0/0 -> NaN in the reference formula, games/race.py:303-306).  All pruned
    kernels must reproduce the literal kernel bit for bit, NaNs included, and the C oracle must agree."""
    from game_level_gan_b200.games import Race, RaceConfig, _tables
    B = 12
    tracks = torch.zeros(B, 128, 2)
    tracks[6:, 40:60, 0] = 0.5                                     # half of them with a bend further on
    noop = torch.zeros((2, B), dtype=torch.int64)
    fwd = torch.ones((2, B), dtype=torch.int64)
    envs = {v: Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant=v)
            for v in ('brute', 'fast', 'warp', 'scan')}
    orc = _c_oracle_for(None, [(60., 4., 40.), (60., 1., 80.)], 1. / 20., 40.)
    st, ct, _ = _tables.heading_tables(128)
    orc.reset(tracks.numpy(), st.numpy(), ct.numpy())
    for env in envs.values():
        env.reset(tracks)
        env.positions[:, 0, 0] = 0.5                               # car 0 on the right wall's line x = +0.5
        env.positions[:, 1, 0] = -0.5                              # car 1 on the left wall's line
        env.positions[::2, :, 1] = 0.                              # every other track: also on the start line
    orc.pos[:, 0, 0] = 0.5
    orc.pos[:, 1, 0] = -0.5
    orc.pos[::2, :, 1] = 0.
    ref = envs['brute']
    saw_nan = False
    for s, a in enumerate([noop, noop, fwd, fwd, noop, fwd]):
        for env in envs.values():
            env.finished()                 # like the reference's loop (train-gan.py:91); arms the 19-wide early-out
        so, ro_ = ref.step(a.cuda())
        co, cr = orc.step(a.numpy())
        assert eq(so, co) and eq(ro_, cr), ('oracle', s)
        saw_nan = saw_nan or bool(torch.isnan(so).any())
        for v in ('fast', 'warp', 'scan'):
            sg, rg = envs[v].step(a.cuda())
            assert eq(sg, so) and eq(rg, ro_), (v, s)
            assert eq(envs[v].alive, ref.alive) and eq(envs[v].positions, ref.positions)
    assert saw_nan or float(so[..., :18].min()) == 0.              # the configuration really is degenerate


def test_wide_and_wild_float_tracks_all_variants():
    """Arbitrary float arcs and widths (walls up to ~2 units long, so almost every wall is "close" and flagged, and
    some cars overflow the candidate queue): every pruned kernel against the literal one, 60 steps."""
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(4242)
    B, T = 192, 60
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.rand((B, 128), generator=g) * 2 - 1
    tracks[:, :, 1] = torch.rand((B, 128), generator=g)
    tracks[: B // 3, :, 0] *= 0.3                                  # a third of them gentle, so that cars survive longer
    acts = torch.randint(0, 9, (T, 2, B), generator=g)
    acts = torch.where(torch.rand((T, 2, B), generator=g) < 0.6, torch.ones_like(acts), acts)
    envs = {v: Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant=v)
            for v in ('brute', 'fast', 'warp', 'scan')}
    s0 = {v: env.reset(tracks)[0] for v, env in envs.items()}
    ref = envs['brute']
    for v in ('fast', 'warp', 'scan'):
        assert eq(s0[v], s0['brute']) and eq(envs[v]._valid_tracks, ref._valid_tracks)
    for s in range(T):
        so, ro_ = ref.step(acts[s].cuda())
        for v in ('fast', 'warp', 'scan'):
            sg, rg = envs[v].step(acts[s].cuda())
            assert eq(sg, so) and eq(rg, ro_), (v, s)
    for v in ('fast', 'warp', 'scan'):
        assert eq(envs[v].positions, ref.positions) and eq(envs[v].scores, ref.scores) and eq(envs[v].winners(), ref.winners())
    assert 0 < int(ref.alive.sum()) or int(ref._valid_tracks.sum()) >= 0


@pytest.mark.parametrize('L', [128, 64, 200, 208, 30, 3])
def test_validity_sign_matrix_equals_the_pair_loop(L):
    """glg_track_validate (every (line, end point) orientation once, bit rows) against glg_track_validate_pairs (the
    literal loop over all pairs of lines, games/race.py:326-334): curly iid-9 tracks (many self-crossings), gentle and
    straight ones (collinear walls: orientation values that are exactly or almost zero), float widths."""
    from game_level_gan_b200 import _lib
    from game_level_gan_b200._lib import check, ptr
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(900 + L)
    B = 600
    tracks = torch.zeros(B, L, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, L), generator=g)]
    tracks[100:200, :, 0] *= 0.25                                              # gentle
    tracks[200:260, :, 0] = 0.                                                 # straight: every wall collinear
    tracks[260:330, :, 0] = torch.where(torch.rand((70, L), generator=g) < 0.8, torch.zeros(70, L), tracks[260:330, :, 0])
    tracks[330:450, :, 0] = torch.rand((120, L), generator=g) * 2 - 1
    tracks[330:450, :, 1] = torch.rand((120, L), generator=g)
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    env.reset(tracks)
    N = L + 2
    a = torch.empty(B, dtype=torch.uint8, device='cuda')
    b = torch.empty(B, dtype=torch.uint8, device='cuda')
    stream = _lib.stream_ptr(env.device)
    check(_lib.lib().glg_track_validate(ptr(env._geom), B, N, ptr(a), stream), 'glg_track_validate')
    check(_lib.lib().glg_track_validate_pairs(ptr(env._geom), B, N, ptr(b), stream), 'glg_track_validate_pairs')
    assert eq(a, b) and eq(a, env._valid_tracks)
    if L >= 30:
        assert 0 < int(a.sum()) < B                                            # both outcomes occur
