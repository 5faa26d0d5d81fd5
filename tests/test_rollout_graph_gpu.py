"""GraphedRollout (one CUDA graph of k x [policy -> env step]) against the reference's per-step loop
(train-gan.py:91-93) run with the same deterministic policies, and against the same loop with the C oracle as the
environment: the environment must end in the same state."""
import pytest
import torch

from tests.helpers import eq

pytestmark = pytest.mark.gpu


class RecurrentArgmaxPolicy(object):
    """LSTMPolicy-like (policies/LSTMPolicy.py:26-41): two LSTM cells + linear head per player, greedy action,
    recurrent state kept in place so the same code runs eagerly and inside a captured graph."""

    def __init__(self, P, B, width, seed, hidden=32):
        g = torch.Generator().manual_seed(seed)
        self.cells = [[torch.nn.LSTMCell(width if i == 0 else hidden, hidden).cuda() for i in range(2)] for _ in range(P)]
        self.heads = [torch.nn.Linear(hidden, 9).cuda() for _ in range(P)]
        for mods in self.cells:
            for m in mods:
                for w in m.parameters():
                    w.data.copy_(torch.randn(w.shape, generator=g) * 0.4)
        for p, m in enumerate(self.heads):
            for w in m.parameters():
                w.data.copy_(torch.randn(w.shape, generator=g) * (0.5 if p else 2.0))
        self.h = [[(torch.zeros(B, hidden, device='cuda'), torch.zeros(B, hidden, device='cuda')) for _ in range(2)] for _ in range(P)]

    def reset(self):
        for per_player in self.h:
            for h, c in per_player:
                h.zero_(); c.zero_()

    def __call__(self, states):
        acts = []
        for p in range(len(self.cells)):
            x = states[p]
            for i, cell in enumerate(self.cells[p]):
                h, c = cell(x, self.h[p][i])
                self.h[p][i][0].copy_(h); self.h[p][i][1].copy_(c)
                x = h
            logits = self.heads[p](x)
            logits[:, 1] += 1.5                      # bias towards "forward" so that cars travel
            acts.append(torch.argmax(logits, dim=-1))
        return torch.stack(acts, 0)


@pytest.mark.parametrize('timeout,k', [(3., 7), (40., 16)])
def test_graphed_rollout_matches_the_reference_loop(timeout, k):
    from game_level_gan_b200.games import GraphedRollout, Race, RaceConfig
    g = torch.Generator().manual_seed(31)
    B = 96
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    with torch.no_grad():
        # the reference's loop
        env = Race(timeout=timeout, cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        pol = RecurrentArgmaxPolicy(2, B, 20, seed=1)
        states, any_valid = env.reset(tracks)
        while any_valid and not env.finished():
            states, rewards = env.step(pol(states))
        # the graphed loop, fresh policy state
        env2 = Race(timeout=timeout, cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        pol2 = RecurrentArgmaxPolicy(2, B, 20, seed=1)
        states2, _ = env2.reset(tracks)
        roll = GraphedRollout(env2, pol2, steps_per_replay=k, on_reset=pol2.reset)
        roll.run(states2)
    assert env2.finished() and env.finished()
    assert env2.steps == env.steps, (env2.steps, env.steps)
    for a, b in ((env.positions, env2.positions), (env.directions, env2.directions), (env.speeds, env2.speeds),
                 (env.alive, env2.alive), (env.finishes, env2.finishes), (env.scores, env2.scores)):
        assert eq(a, b)
    assert eq(env.winners(), env2.winners())
    if timeout > 10:
        assert int(env.finishes.sum()) + int((~env.alive).sum()) > 0
    # the same episode with the C ORACLE as the environment (pinned to the reference's fixtures) and a third instance
    # of the policy: the closed loop ends in the same state, so the two CUDA loops above are reference-equivalent
    import ctypes
    from game_level_gan_b200.games import _tables
    from oracle import c_oracle as co
    cpr = co.RaceParams()
    ctypes.memmove(ctypes.byref(cpr), ctypes.byref(_tables.race_params(RaceConfig.cars, 1. / 20., timeout, 18, 10.)), ctypes.sizeof(cpr))
    orc = co.CRace(cpr)
    st_, ct_, _ = _tables.heading_tables(128)
    with torch.no_grad():
        pol3 = RecurrentArgmaxPolicy(2, B, 20, seed=1)
        so, any_valid = orc.reset(tracks.numpy(), st_.numpy(), ct_.numpy())
        while any_valid and not orc.finished():
            so, _ = orc.step(pol3(torch.from_numpy(so).cuda()).cpu().numpy())
    assert orc.steps == env.steps
    assert eq(env.positions, orc.pos) and eq(env.directions, orc.dir) and eq(env.speeds, orc.speed)
    assert eq(env._alive, orc.alive) and eq(env._finishes, orc.finishes) and eq(env.scores, orc.scores)
    assert eq(env.winners(), orc.winners())


def test_host_stepper_matches_step():
    """HostStepper (one CUDA graph per step, host buffers) against Race.step, through deaths, the time limit and
    the 19-wide early-out, and across a rewind."""
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(8)
    B, T = 200, 70
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)]
    acts = torch.randint(0, 9, (T, 2, B), generator=g)
    acts[:, :, ::3] = 1
    a = Race(timeout=3., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    b = Race(timeout=3., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    a.reset(tracks)
    b.reset(tracks)
    hs = b.host_stepper()
    snap = b.snapshot()
    for s in range(5):
        hs.step(acts[s])
    b.restore(snap)
    for s in range(T):
        sa, ra = a.step(acts[s].cuda())
        sb, rb = hs.step(acts[s] if s % 2 else acts[s].numpy())
        assert sb.device.type == 'cpu' and eq(sa, sb) and eq(ra, rb), 'step %d' % s
        assert a.finished() == b.finished() and a.steps == b.steps
    assert eq(a.positions, b.positions) and eq(a.scores, b.scores) and eq(a.winners(), b.winners())
    assert a.finished()
    # everybody dead: tiny batch of cars driving into the wall
    c = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    d = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
    tr = tracks[:4]
    c.reset(tr); d.reset(tr)
    hd = d.host_stepper()
    right = torch.full((2, 4), 4, dtype=torch.int64)          # forward-right until the wall
    for s in range(400):
        c.finished()                       # `step` itself never blocks: the early-out needs the loop's own question
        sc, rc = c.step(right.cuda())
        sd, rd = hd.step(right)
        assert sc.shape == sd.shape and eq(sc, sd) and eq(rc, rd), 'step %d' % s
        if sc.size(-1) == 19:
            break
    assert sc.size(-1) == 19 and c.finished() and d.finished()
    # a stepper is bound to its episode
    from game_level_gan_b200._lib import GlgError
    d.reset(tr)
    with pytest.raises(GlgError):
        hd.step(right)


def test_graphed_rollout_history_and_state_after_a_time_limit_end():
    """An episode that ends by the time limit with cars alive: the graph's replays past the limit are no-ops, the
    environment afterwards answers like after the reference's loop (somebody IS alive, so a further `step` is a normal
    20-wide step, games/race.py:353-356), and with `log_history` the recorded board's history equals the loop's."""
    from game_level_gan_b200.games import GraphedRollout, Race, RaceConfig
    g = torch.Generator().manual_seed(5)
    B = 64
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)] * 0.3
    with torch.no_grad():
        env = Race(timeout=2., cars=RaceConfig.cars, framerate=1. / 20., log_history=True)
        env.record(3)
        pol = RecurrentArgmaxPolicy(2, B, 20, seed=4)
        states, any_valid = env.reset(tracks)
        while any_valid and not env.finished():
            states, rewards = env.step(pol(states))
        env2 = Race(timeout=2., cars=RaceConfig.cars, framerate=1. / 20., log_history=True)
        env2.record(3)
        pol2 = RecurrentArgmaxPolicy(2, B, 20, seed=4)
        states2, _ = env2.reset(tracks)
        GraphedRollout(env2, pol2, steps_per_replay=7, on_reset=pol2.reset).run(states2)
        assert env.steps == env2.steps == env.steps_limit + 1 and int(env.alive.sum()) > 0
        assert env2._any_alive() and env._any_alive()
        h1, h2 = env.history, env2.history
        assert len(h1) == len(h2) == env.steps and h1 == h2
        noop = torch.zeros((2, B), dtype=torch.int64, device='cuda')
        s1, r1 = env.step(noop)
        s2, r2 = env2.step(noop)
        assert s1.shape == s2.shape == (2, B, 20) and eq(s1, s2) and eq(r1, r2)


def test_host_stepper_records_history():
    from game_level_gan_b200.games import Race, RaceConfig
    g = torch.Generator().manual_seed(9)
    tracks = torch.zeros(8, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (8, 128), generator=g)]
    acts = torch.randint(0, 9, (30, 2, 8), generator=g)
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=True)
    b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=True)
    for e in (a, b):
        e.record(2)
        e.reset(tracks)
    hs = b.host_stepper()
    for s in range(30):
        a.finished()
        a.step(acts[s].cuda())
        hs.step(acts[s])
    assert len(a.history) == 31 and a.history == b.history


def _stock_agents(seed, P=2):
    """The reference's own PPOAgent + LSTMPolicy (agents/PPOAgent.py, policies/LSTMPolicy.py) from baseline/_ref (made by
    tools/make_baseline_ref.py; travels with the working tree, absent in a bare checkout -> skip)."""
    import os
    import sys
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    if not os.path.isdir(os.path.join(root, 'agents')):
        pytest.skip('baseline/_ref/agents not present (python tools/make_baseline_ref.py)')
    if root not in sys.path:
        sys.path.insert(0, root)
    from agents import PPOAgent
    from policies import LSTMPolicy
    torch.manual_seed(seed)
    agents = [PPOAgent(9, LSTMPolicy(20, 9)) for _ in range(P)]
    for a in agents:
        with torch.no_grad():
            a.network.policy.bias[1] += 1.0          # bias towards "forward" so that cars travel
        a.reset()                                    # old_network <- network
    return agents


def test_graphed_rollout_with_the_reference_ppo_agents():
    """train-gan.py:86-96 with the STOCK agents: the reference's loop (`a.act(s, training=False)` per agent and step,
    greedy so that both runs are deterministic) against GraphedRollout driving the same agents through
    `capture_safe_agents` - same final environment, same step count, same winners."""
    from game_level_gan_b200.games import GraphedRollout, Race, RaceConfig, capture_safe_agents
    g = torch.Generator().manual_seed(77)
    B = 128
    tracks = torch.zeros(B, 128, 2)
    tracks[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (B, 128), generator=g)] * 0.5
    with torch.no_grad():
        env = Race(timeout=6., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        agents = _stock_agents(3)
        states, any_valid = env.reset(tracks)
        while any_valid and not env.finished():
            actions = torch.stack([a.act(s, deterministic=True, training=False) for a, s in zip(agents, states)], dim=0)
            states, rewards = env.step(actions)
        for a in agents:
            a.reset()
        env2 = Race(timeout=6., cars=RaceConfig.cars, framerate=1. / 20., log_history=False)
        agents2 = _stock_agents(3)
        act, on_reset = capture_safe_agents(agents2, deterministic=True)
        states2, _ = env2.reset(tracks)
        roll = GraphedRollout(env2, act, steps_per_replay=10, on_reset=on_reset)
        roll.run(states2)
        assert env.steps == env2.steps and env.steps > 20
        for a, b in ((env.positions, env2.positions), (env.directions, env2.directions), (env.speeds, env2.speeds),
                     (env.alive, env2.alive), (env.finishes, env2.finishes), (env.scores, env2.scores)):
            assert eq(a, b)
        assert eq(env.winners(), env2.winners())
        assert float(env.positions[:, :, 1].max()) > 1.0                    # the cars did drive
        # a second episode with the same graph (agents reset, new states): still the loop's result
        states2, _ = env2.reset(tracks)
        roll2 = GraphedRollout(env2, act, steps_per_replay=10, on_reset=on_reset)
        roll2.run(states2)
        assert env2.steps == env.steps and eq(env.positions, env2.positions)
