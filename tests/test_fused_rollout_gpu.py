"""The persistent rollout kernel (glg_race_rollout, GLG_ROLLOUT_FUSED: one launch plays T steps, track records in
shared memory, car state in registers) against the reference-generated fixtures, the literal step kernel and the
per-step production kernel - bit for bit, in chunks of different lengths, with and without keep_all."""
import pytest
import torch

from tests.helpers import RACE_CASES, eq, load_case, nmismatch, t

pytestmark = pytest.mark.gpu


def _env(case, variant='fast', log_history=False):
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in case['cars'].tolist()]
    return Race(timeout=float(case['timeout']), cars=cars, framerate=float(case['framerate']),
                log_history=log_history, variant=variant)


@pytest.mark.parametrize('name', RACE_CASES)
def test_fused_rollout_reproduces_reference_fixture(name):
    """Every reference fixture replayed through fused rollouts (chunks of 1, 7, 2 and the rest): all observations,
    rewards and, at every chunk end, the whole car state equal the reference's (games/race.py:340-500)."""
    c = load_case(name)
    geometry = (t(c['centre']), t(c['left']), t(c['right'])) if name == 'floatw' else None
    env = _env(c)
    states, any_valid = env.reset(t(c['tracks']), geometry=geometry)
    assert nmismatch(states, c['states'][0]) == 0
    acts = t(c['actions']).cuda()
    # the rollout has no "nobody alive" early-out: stop where the reference switches to 19-wide outputs
    T = 0
    while T < acts.size(0) and int(c['widths'][T + 1]) == 20:
        T += 1
    assert T >= 10
    s = 0
    for n in (1, 7, 2, T):
        n = min(n, T - s)
        if n <= 0:
            break
        st, rw = env.rollout(acts[s:s + n], keep_all=True, mode='fused')
        assert nmismatch(st, c['states'][s + 1:s + n + 1]) == 0, (name, s, n)
        assert nmismatch(rw, c['rewards'][s:s + n]) == 0, (name, s, n)
        s += n
        for k, v in (('pos', env.positions), ('dir', env.directions), ('speed', env.speeds),
                     ('alive', env.alive), ('finishes', env.finishes), ('scores', env.scores)):
            assert nmismatch(v, c[k][s]) == 0, (name, k, s)
        assert env.finished() == bool(c['finished'][s]) and env.steps == s + 1
    if T == acts.size(0):
        assert eq(env.winners(), c['winners'])


def _iid9(B, seed):
    g = torch.Generator().manual_seed(seed)
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    return tracks, g


def test_fused_rollout_unpruned_fallbacks_and_degenerate_cars():
    """Cars that take the in-kernel brute-force paths (heading norm far from 1, far from the origin), cars exactly
    on a wall's line (all-rays candidates, queue overflow, NaN readings): fused rollout of the production kernel
    against per-step calls of the LITERAL kernel."""
    from game_level_gan_b200.games import Race, RaceConfig
    tracks, g = _iid9(96, 31)
    tracks[:12] = 0.
    tracks[6:12, 40:60, 0] = 0.5
    T = 30
    acts = torch.randint(0, 9, (T, 2, 96), generator=g)
    acts = torch.where(torch.rand((T, 2, 96), generator=g) < 0.5, torch.ones_like(acts), acts)
    acts[:2, :, :12] = 0
    envs = {v: Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant=v)
            for v in ('brute', 'fast')}
    for env in envs.values():
        env.reset(tracks)
        env.positions[:12, 0, 0] = 0.5                             # on the right wall's line
        env.positions[:12, 1, 0] = -0.5
        env.positions[:12:2, :, 1] = 0.                            # and on the start line
        env.directions[12::3] *= 1.6                               # |d|^2 = 2.56
        env.positions[13::5] += 300.
        env.directions[14::7] *= 0.5
    ref = [envs['brute'].step(acts[s].cuda()) for s in range(T)]
    st, rw = envs['fast'].rollout(acts.cuda(), keep_all=True, mode='fused')
    assert eq(st, torch.stack([s for s, _ in ref])) and eq(rw, torch.stack([r for _, r in ref]))
    a, b = envs['brute'], envs['fast']
    for x, y in ((a.positions, b.positions), (a.directions, b.directions), (a.speeds, b.speeds), (a.alive, b.alive),
                 (a.finishes, b.finishes), (a.scores, b.scores)):
        assert eq(x, y)
    assert bool(torch.isnan(st).any()) or float(st[..., :18].min()) == 0.


def test_fused_rollout_wide_float_tracks_and_last_only():
    """Float-width tracks (most walls flagged, queue overflows); keep_all=False returns the last step only and
    leaves the same state."""
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(4243)
    B, T = 160, 50
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.rand((B, 128), generator=g) * 2 - 1
    tracks[:, :, 1] = torch.rand((B, 128), generator=g)
    tracks[: B // 3, :, 0] *= 0.3
    acts = torch.randint(0, 9, (T, 2, B), generator=g)
    acts = torch.where(torch.rand((T, 2, B), generator=g) < 0.6, torch.ones_like(acts), acts)
    mk = lambda v: Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant=v)
    a, b, c = mk('brute'), mk('fast'), mk('fast')
    for env in (a, b, c):
        env.reset(tracks)
    ref = [a.step(acts[s].cuda()) for s in range(T)]
    st, rw = b.rollout(acts.cuda(), keep_all=True, mode='fused')
    assert eq(st, torch.stack([s for s, _ in ref])) and eq(rw, torch.stack([r for _, r in ref]))
    s_last, r_last = c.rollout(acts.cuda(), keep_all=False, mode='fused')
    assert eq(s_last, ref[-1][0]) and eq(r_last, ref[-1][1])
    for x, y, z in ((a.positions, b.positions, c.positions), (a.scores, b.scores, c.scores), (a.alive, b.alive, c.alive)):
        assert eq(x, y) and eq(x, z)
    assert eq(a.winners(), b.winners()) and a.finished() == b.finished() == c.finished()


def test_fused_rollout_history_and_preallocated_outputs():
    """`history` of the recorded board is the same whether the episode is stepped or rolled out (games/race.py:492-494);
    `out=` buffers are used as given; a plan refuses to outlive its episode."""
    from game_level_gan_b200._lib import GlgError
    c = load_case('predef')
    a, b = _env(c, log_history=True), _env(c, log_history=True)
    a.record(3)
    b.record(3)
    a.reset(t(c['tracks']))
    b.reset(t(c['tracks']))
    acts = t(c['actions'][:25]).cuda()
    for s in range(25):
        a.step(acts[s])
    out = (torch.zeros((25, 2, 12, 20), device='cuda'), torch.zeros((25, 2, 12), device='cuda'))
    plan = b.rollout_plan(acts, keep_all=True, mode='fused', out=out)
    st, rw = plan.run()
    assert st.data_ptr() == out[0].data_ptr() and plan.launches == 1
    assert nmismatch(st, c['states'][1:26]) == 0
    ha, hb = a.history, b.history
    assert len(ha) == len(hb) == 26 and ha == hb
    b.reset(t(c['tracks']))
    with pytest.raises(GlgError):
        plan.run()
    with pytest.raises(ValueError):
        b.rollout(acts[:, :1], mode='fused')


def test_step_does_not_block_unless_asked():
    """`step` enqueues without a host synchronisation; the reference's 19-wide "nobody alive" early-out
    (games/race.py:353-356) applies once `finished()` has told the host that everybody is gone."""
    c = load_case('p1_crash')
    env = _env(c)
    env.reset(t(c['tracks']))
    acts = t(c['actions']).cuda()
    widths = []
    for s in range(acts.size(0)):
        st, _ = env.step(acts[s])
        widths.append(st.size(-1))
        env.finished()
    assert widths == [int(w) for w in c['widths'][1:]]
    env2 = _env(c)
    env2.reset(t(c['tracks']))
    for s in range(acts.size(0)):
        st, rw = env2.step(acts[s])                     # never asks: always the kernel, always 20 wide
        assert st.size(-1) == 20
        assert nmismatch(rw, c['rewards'][s]) == 0
    assert eq(env2.positions, env.positions) and eq(env2.scores, env.scores) and env2.steps == env.steps


def test_sharded_winner_stats_equal_the_unsharded_ones():
    """shard the trial-major batch by BOARD (all trials of a board on one rank, train-gan.py:84, 103-104),
    play each shard separately (virtual ranks on one GPU) and concatenate the per-board winner statistics: equal to
    the statistics of the unsharded batch, and to one_hot(winners + 1).view(trials, -1, P + 1).mean(0)."""
    from game_level_gan_b200 import dist as gdist
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)]]
    trials, boards, T, P = 3, 50, 120, 4
    base, g = _iid9(boards, 8)
    tracks = base.repeat(trials, 1, 1)                                  # trial-major
    acts = torch.randint(0, 9, (T, P, trials * boards), generator=g)
    acts = torch.where(torch.rand(acts.shape, generator=g) < 0.6, torch.ones_like(acts), acts)
    whole = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
    whole.reset(tracks)
    whole.rollout(acts.cuda(), mode='fused')
    ref = whole.winner_stats(trials)
    w = whole.winners().cpu()
    assert eq(ref, torch.nn.functional.one_hot(w + 1, P + 1).view(trials, -1, P + 1).float().mean(0))
    world = 3
    parts = []
    for rank in range(world):
        local, (lo, hi) = gdist.shard_trial_major(tracks, trials, rank, world)
        a_local = acts.view(T, P, trials, boards)[:, :, :, lo:hi].reshape(T, P, -1)
        env = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False)
        env.reset(local)
        env.rollout(a_local.cuda(), mode='fused')
        parts.append(env.winner_stats(trials))
        assert parts[-1].shape == (hi - lo, P + 1)
    assert eq(torch.cat(parts), ref)
    assert float(ref[:, 1:].sum()) > 0                                  # some boards do have winners


@pytest.mark.parametrize('L,P', [(64, 2), (200, 2), (254, 3), (30, 4), (10, 2)])
def test_fused_rollout_other_track_lengths(L, P):
    """Track lengths other than the reference's 128 segments take the generic instantiations of the fused kernel
    (run-time shared-memory layout and trip counts; L = 10 is shorter than the arg-min window): against per-step calls
    of the literal kernel."""
    from game_level_gan_b200.games import Race, RaceCar
    cars = [RaceCar(*c) for c in [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)][:P]]
    g = torch.Generator().manual_seed(500 + L)
    B, T = 77, 40
    tracks = torch.zeros(B, L, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, L), generator=g)]
    tracks[B // 2:, :, 1] = torch.rand((B - B // 2, L), generator=g)
    acts = torch.randint(0, 9, (T, P, B), generator=g)
    acts = torch.where(torch.rand((T, P, B), generator=g) < 0.6, torch.ones_like(acts), acts)
    a = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, variant='brute')
    b = Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, variant='fast')
    sa0, _ = a.reset(tracks)
    sb0, _ = b.reset(tracks)
    assert eq(sa0, sb0)
    ref = [a.step(acts[s].cuda()) for s in range(T)]
    plan = b.rollout_plan(acts.cuda(), keep_all=True, mode='fused')
    assert plan.launches == 1
    st, rw = plan.run()
    assert eq(st, torch.stack([s for s, _ in ref])) and eq(rw, torch.stack([r for _, r in ref]))
    for x, y in ((a.positions, b.positions), (a.directions, b.directions), (a.speeds, b.speeds), (a.alive, b.alive),
                 (a.finishes, b.finishes), (a.scores, b.scores)):
        assert eq(x, y)
    assert eq(a.winners(), b.winners())


@pytest.mark.parametrize('T,chunk', [(40, 25), (37, 5), (12, 40)])
def test_host_rollout_equals_device_rollout(T, chunk):
    """`Race.host_rollout`: pinned host action tape in, pinned host observations / rewards out, chunks pipelined over
    copy-in / compute / copy-out streams - same results as `Race.rollout(keep_all=True)` and as per-step calls of the
    literal kernel, twice in a row (the staging buffers are reused) and after a rewind."""
    from game_level_gan_b200.games import Race, RaceConfig
    tracks, g = _iid9(300, 91)
    acts = torch.randint(0, 9, (T, 2, 300), generator=g)
    acts = torch.where(torch.rand((T, 2, 300), generator=g) < 0.6, torch.ones_like(acts), acts)
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant='brute')
    b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    a.reset(tracks)
    b.reset(tracks)
    ref = [a.step(acts[s].cuda()) for s in range(T)]
    ref_s, ref_r = torch.stack([s for s, _ in ref]), torch.stack([r for _, r in ref])
    snap = b.snapshot()
    hr = b.host_rollout(T, chunk=chunk)
    assert hr.launches == len(hr.bounds) and hr.bounds[0][0] == 0 and hr.bounds[-1][1] == T
    assert all(a[1] == b[0] for a, b in zip(hr.bounds[:-1], hr.bounds[1:]))            # the chunks tile [0, T)
    assert hr.bounds[0][1] == min(T, max(1, chunk // 5)) and all(hi - lo <= chunk for lo, hi in hr.bounds)
    assert hr.h2d_bytes == T * 2 * 300 * 8 and hr.d2h_bytes == T * 2 * 300 * 21 * 4
    for rep in range(2):
        st, rw = hr.run(acts if rep == 0 else acts.numpy())
        assert not st.is_cuda and st.is_pinned() and st.shape == (T, 2, 300, 20)
        assert eq(st, ref_s) and eq(rw, ref_r), rep
        assert b.steps == T + 1 and eq(b.positions, a.positions) and eq(b.alive, a.alive) and eq(b.scores, a.scores)
        st.zero_()
        rw.zero_()
        b.restore(snap)
    with pytest.raises(ValueError):
        hr.run(acts[:-1])
    b.reset(tracks)
    with pytest.raises(Exception):
        hr.run(acts)


def test_host_rollouts_submitted_alternately_advance_one_episode():
    """Two HostRollouts of one environment used in turn (`submit` the next call before `wait`ing for the previous one):
    the kernels of all calls run in order, so four 15-step calls equal 60 per-step calls of the literal kernel."""
    from game_level_gan_b200.games import Race, RaceConfig
    tracks, g = _iid9(200, 92)
    T, calls = 15, 4
    acts = torch.randint(0, 9, (calls * T, 2, 200), generator=g)
    acts = torch.where(torch.rand(acts.shape, generator=g) < 0.6, torch.ones_like(acts), acts)
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, variant='brute')
    b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    a.reset(tracks)
    b.reset(tracks)
    ref = [a.step(acts[s].cuda()) for s in range(calls * T)]
    ref_s, ref_r = torch.stack([s for s, _ in ref]), torch.stack([r for _, r in ref])
    hrs = [b.host_rollout(T, chunk=6), b.host_rollout(T, chunk=6)]
    got_s, got_r, pending = [], [], None
    for i in range(calls):
        h = hrs[i % 2]
        h.submit(acts[i * T:(i + 1) * T])
        if pending is not None:
            st, rw = pending.wait()
            got_s.append(st.clone())
            got_r.append(rw.clone())
        pending = h
    st, rw = pending.wait()
    got_s.append(st.clone())
    got_r.append(rw.clone())
    assert eq(torch.cat(got_s), ref_s) and eq(torch.cat(got_r), ref_r)
    assert b.steps == a.steps and eq(b.positions, a.positions) and eq(b.scores, a.scores)
    with pytest.raises(Exception):
        hrs[0].wait()                                  # nothing pending
    hrs[0].submit(acts[:T])
    with pytest.raises(Exception):
        hrs[0].submit(acts[:T])                        # not waited for
    hrs[0].wait()
