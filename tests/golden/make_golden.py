"""Generate the golden fixtures in this directory from the REAL reference implementation.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's C++ helper cannot be built in this image (Boost.Geometry is absent), so the
reference itself falls back to its torch path (``IMPL_GPU``, games/race.py:89-101).  To skip the
doomed ~6 s JIT build, ``torch.utils.cpp_extension.load`` is replaced by a stub raising the same
``RuntimeError`` the failed build raises - the reference then takes exactly the branch it takes
when run stock in this image.  Everything runs on CPU (the canonical oracle).

Each ``race_*.npz`` holds: inputs (tracks, actions, car parameters, timeout), the geometry the
reference built, validity, and for every step all outputs and the full public state.
``kat_*.npz`` hold known-answer tables for the static geometry predicates.
"""
import contextlib
import io
import math
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def import_reference():
    from torch.utils import cpp_extension as ext

    def _no_boost(*a, **k):
        raise RuntimeError('game_helpers.cpp: boost/geometry.hpp: No such file or directory')

    ext.load = _no_boost
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    with contextlib.redirect_stdout(io.StringIO()):
        import games  # noqa: F401  (builds the module-level race_game)
    os.chdir(cwd)
    return games


def make_env(games, cars, timeout=40., framerate=1. / 20.):
    with contextlib.redirect_stdout(io.StringIO()):
        env = games.Race(timeout=timeout, framerate=framerate,
                         cars=[games.RaceCar(*c) for c in cars],
                         log_history=False, device=torch.device('cpu'))
    assert env._impl_version == games.Race.IMPL_GPU
    return env


def snapshot(env):
    return dict(pos=env.positions.clone(), dir=env.directions.clone(), speed=env.speeds.clone(),
                alive=env.alive.clone(), finishes=env.finishes.clone(), scores=env.scores.clone())


def run_case(games, name, tracks, cars, actions=None, policy=None, steps=None, timeout=40.,
             framerate=1. / 20.):
    """actions [T,P,B] int64, or policy(states[P,B,W], t) -> [P,B] actions."""
    env = make_env(games, cars, timeout, framerate)
    states, any_valid = env.reset(tracks.clone())
    rec = dict(states=[states.contiguous().clone()], rewards=[], widths=[states.size(-1)],
               finished=[bool(env.finished())])
    snaps = [snapshot(env)]
    acts = []
    T = actions.size(0) if actions is not None else steps
    for t in range(T):
        a = actions[t] if actions is not None else policy(states, t)
        acts.append(a.clone())
        states, rewards = env.step(a.clone())
        w = states.size(-1)
        st = torch.zeros(states.size(0), states.size(1), 20)
        st[:, :, :w] = states
        rec['states'].append(st)
        rec['widths'].append(w)
        rec['rewards'].append(rewards.contiguous().clone())
        rec['finished'].append(bool(env.finished()))
        snaps.append(snapshot(env))
    out = dict(
        tracks=tracks.numpy(), cars=np.array(cars, dtype=np.float64), timeout=np.float64(timeout),
        framerate=np.float64(framerate), actions=torch.stack(acts).numpy(),
        centre=env.segments.numpy(), left=env.left_vecs.numpy(), right=env.right_vecs.numpy(),
        valid=env.valid.numpy(), any_valid=np.bool_(any_valid),
        states=torch.stack(rec['states']).numpy(), widths=np.array(rec['widths']),
        rewards=torch.stack(rec['rewards']).numpy(), finished=np.array(rec['finished']),
        winners=env.winners().numpy(), steps_limit=np.int64(env.steps_limit),
    )
    for k in snaps[0]:
        out[k] = torch.stack([s[k] for s in snaps]).numpy()
    path = os.path.join(HERE, 'race_%s.npz' % name)
    np.savez_compressed(path, **out)
    alive_end = int(out['alive'][-1].sum())
    print('%-14s B=%d P=%d T=%d valid=%d finishes=%d alive_end=%d  %.0f KB' % (
        name, tracks.size(0), len(cars), T, int(out['valid'].sum()) // len(cars),
        int(out['finishes'][-1].sum()), alive_end, os.path.getsize(path) / 1024))


def biased_actions(T, P, B, p_forward, gen):
    a = torch.randint(0, 9, (T, P, B), generator=gen)
    fwd = torch.rand((T, P, B), generator=gen) < p_forward
    return torch.where(fwd, torch.ones_like(a), a)


def load_agents(n_players):
    """Shipped LSTM agents (learned/agent_{0,1}_0.pt) as greedy-ish action sources."""
    from policies import LSTMPolicy
    nets = []
    for i in range(n_players):
        net = LSTMPolicy(20, 9)
        sd = torch.load(os.path.join(REF, 'learned', 'agent_%d_0.pt' % (i % 2)),
                        map_location='cpu', weights_only=False)
        net.load_state_dict(sd['network'])
        net.eval()
        nets.append(net)
    return nets


def main():
    games = import_reference()
    default = [(60., 4., 40.), (60., 1., 80.)]
    four = default + [(80., 2., 60.), (50., 3., 50.)]
    space = torch.linspace(-1., 1., 9)
    g = torch.Generator().manual_seed(20261018)

    # 1. predefined tracks + mirror (train-gan.py:82-83), forward-biased random actions
    random.seed(0)
    pre = games.predefined_tracks().cpu()
    pm = torch.cat((pre, -pre), dim=0)
    run_case(games, 'predef', pm, default, actions=biased_actions(150, 2, 12, 0.6, g))

    # 2. iid 9-level arcs (generator-like), uniform random actions
    t = torch.zeros(16, 128, 2)
    t[:, :, 0] = space[torch.randint(0, 9, (16, 128), generator=g)]
    run_case(games, 'iid9', t, default, actions=torch.randint(0, 9, (120, 2, 16), generator=g))

    # 3. float arcs and non-zero widths
    t = torch.zeros(8, 128, 2)
    t[:, :, 0] = torch.rand((8, 128), generator=g) * 1.6 - 0.8
    t[:, :, 1] = torch.rand((8, 128), generator=g)
    run_case(games, 'floatw', t, default, actions=biased_actions(100, 2, 8, 0.7, g))

    # 4. self-intersecting (invalid) loops mixed with valid tracks
    t = torch.zeros(6, 128, 2)
    t[0, :, 0] = 1.
    t[1, :60, 0] = 1.
    t[2, :, 0] = 0.5
    t[3, 20:75, 0] = -1.
    t[4, :, 0] = 0.
    t[5, 40:60, 0] = 1.
    run_case(games, 'loops', t, default, actions=biased_actions(40, 2, 6, 0.6, g))

    # 5. one player: everybody crashes, then keep stepping (19-wide early-out, race.py:353-356)
    t = torch.zeros(3, 128, 2)
    t[1, :, 0] = space[torch.randint(0, 9, (128,), generator=g)]
    a = torch.full((60, 1, 3), 7, dtype=torch.int64)   # forward-left until the wall
    run_case(games, 'p1_crash', t, default[:1], actions=a)

    # 6. four players, short track (L=40), short timeout -> timeout ending and argmax winners
    t = torch.zeros(6, 40, 2)
    t[:, :, 0] = space[torch.randint(2, 7, (6, 40), generator=g)]
    run_case(games, 'p4_short', t, four, actions=biased_actions(70, 4, 6, 0.75, g), timeout=3.)

    # 7. shipped agents drive to the finish (finishes, scores=steps, argmin winners)
    nets = load_agents(2)
    hidden = [None, None]
    ag = torch.Generator().manual_seed(7)

    def policy(states, step):
        acts = []
        with torch.no_grad():
            for i, net in enumerate(nets):
                if states.size(-1) != 20:
                    acts.append(torch.zeros(states.size(1), dtype=torch.int64))
                    continue
                pol, _ = net(states[i].contiguous())
                u = torch.rand(pol.shape, generator=ag)
                gum = -torch.log(-torch.log(u + 1e-8) + 1e-8)
                acts.append(torch.argmax(pol + gum, dim=-1))
        return torch.stack(acts)

    random.seed(3)
    pre = games.predefined_tracks().cpu()
    for net in nets:
        net.reset(True) if hasattr(net, 'reset') else None
    run_case(games, 'agents', torch.cat((pre, -pre), dim=0)[:8], default, policy=policy, steps=360)

    # ---- known-answer tables for the static predicates ---------------------------------
    R = games.Race
    # integer-lattice segments: exact zeros, touching and collinear configurations
    segs = torch.randint(-3, 4, (40, 12, 4), generator=g).float()
    probes = torch.randint(-3, 4, (40, 5, 4), generator=g).float()
    cross = R._segment_collisions(segs, probes)
    hit, start_on = R._segment_collisions(segs, probes, special=True)
    # rays on the lattice (includes start-on-wall, parallel/collinear -> NaN, behind -> inf)
    rays = torch.randint(-3, 4, (40, 6, 4), generator=g).float()
    dist_lat = R._smallest_distance(segs, rays)
    # generic float configuration
    fs = torch.randn((30, 20, 4), generator=g)
    fr = torch.cat((torch.randn((30, 7, 2), generator=g) * 0.3,
                    torch.nn.functional.normalize(torch.randn((30, 7, 2), generator=g), dim=-1)), -1)
    dist_f = R._smallest_distance(fs, fr)
    env = make_env(games, default)
    lat_lines = torch.randint(-4, 5, (50, 7, 4), generator=g).float()
    ok = env._is_correct(lat_lines)
    np.savez_compressed(os.path.join(HERE, 'kat_predicates.npz'),
                        segs=segs.numpy(), probes=probes.numpy(), cross=cross.numpy(),
                        hit=hit.numpy(), start_on=start_on.numpy(), rays=rays.numpy(),
                        dist_lat=dist_lat.numpy(), fsegs=fs.numpy(), frays=fr.numpy(),
                        dist_f=dist_f.numpy(), lat_lines=lat_lines.numpy(), lat_ok=ok.numpy())
    print('kat_predicates  nan=%d inf=%d zero=%d' % (
        int(torch.isnan(dist_lat).sum()), int(torch.isinf(dist_lat).sum()), int((dist_lat == 0).sum())))

    # rotation / sensor-angle tables the kernels take from the host
    ang = torch.tensor([c[2] * math.pi / 180. for c in four], dtype=torch.float32)
    v = torch.randn((64, 2), generator=g)
    a = (torch.rand((64,), generator=g) - 0.5) * 0.2
    np.savez_compressed(os.path.join(HERE, 'kat_rotate.npz'), vecs=v.numpy(), angles=a.numpy(),
                        out=R._rotate_vecs(v, a).numpy(), car_angle=ang.numpy())


def import_reference_pacman():
    """games/pacman.py indexes with lists of arrays (`grid[list(idx.T)]`), which numpy >= 1.23 rejects
    (SURVEY.md section 2, row 5).  The only change made here is list(...) -> tuple(...) at those five
    places; the source is patched in memory, nothing is written to disk."""
    import types
    src = open(os.path.join(REF, 'games', 'pacman.py')).read()
    src = src.replace('from .environment import MultiEnvironment', 'MultiEnvironment = object')
    for a, b in (('self.grid[list(pos_ins.T)]', 'self.grid[tuple(pos_ins.T)]'),
                 ('positions = list(np.transpose(self.players + np.array([0, 0, 0, self.fields])))',
                  'positions = tuple(np.transpose(self.players + np.array([0, 0, 0, self.fields])))'),
                 ('small = list(small.T)', 'small = tuple(small.T)'),
                 ('large = list(large.T)', 'large = tuple(large.T)')):
        assert a in src
        src = src.replace(a, b)
    mod = types.ModuleType('ref_pacman')
    exec(compile(src, 'ref_pacman.py', 'exec'), mod.__dict__)
    return mod.Pacman


def random_levels(B, H, W, P, rng, corners=True):
    """Levels like generators/pacman_generator.py:56-63: iid fields, players in the corners."""
    fields = rng.choice(4, size=(B, H, W), p=[0.4, 0.5, 0.07, 0.03])
    board = np.zeros((B, H, W, 4 + P), dtype=np.int32)
    np.put_along_axis(board[..., :4], fields[..., None], 1, axis=-1)
    spots = [(0, 0), (H - 1, W - 1), (0, W - 1), (H - 1, 0)]
    for b in range(B):
        for p in range(P):
            x, y = spots[p] if corners else (rng.integers(0, H), rng.integers(0, W))
            board[b, x, y, :4] = 0
            board[b, x, y, 0] = 1          # players start on an empty cell
            board[b, x, y, 4 + p] = 1
    return board


def run_pacman_case(Pacman, name, board, P, T, rng, actions=None):
    B, H, W = board.shape[:3]
    env = Pacman((H, W), P, batch_size=B)
    obs = env.reset(board.copy())
    grids, rewards, acts = [env.grid.copy()], [], []
    first_obs = np.stack(obs)
    for t in range(T):
        a = actions[t] if actions is not None else rng.integers(0, 5, size=(B, P)).astype(np.int32)
        acts.append(a)
        obs, rew = env.step(a)
        grids.append(env.grid.copy())
        rewards.append(np.stack(rew))
    np.savez_compressed(os.path.join(HERE, 'pacman_%s.npz' % name), board=board, actions=np.stack(acts),
                        grids=np.stack(grids).astype(np.int8), rewards=np.stack(rewards),
                        first_obs=first_obs, last_obs=np.stack(obs), players=env.players)
    print('pacman_%-10s B=%d %dx%d P=%d T=%d reward_sum=%.2f' % (name, B, H, W, P, T, float(np.sum(rewards))))


def main_pacman():
    Pacman = import_reference_pacman()
    rng = np.random.default_rng(20261018)
    run_pacman_case(Pacman, 'p2', random_levels(6, 15, 15, 2, rng), 2, 40, rng)
    run_pacman_case(Pacman, 'p3', random_levels(4, 9, 11, 3, rng), 3, 30, rng)          # np.where order quirk
    run_pacman_case(Pacman, 'p4_random', random_levels(3, 7, 7, 4, rng, corners=False), 4, 30, rng)
    # single board where nobody can move on some steps (batch-wide skip, pacman.py:77)
    board, size, P = Pacman.from_str('#####\n#1s2#\n#####')
    acts = np.array([[[1, 1]], [[2, 2]], [[4, 3]], [[1, 2]], [[3, 4]], [[0, 0]]], dtype=np.int32)
    run_pacman_case(Pacman, 'walled', board, P, len(acts), rng, actions=acts)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'pacman':
        main_pacman()
    else:
        main()
        main_pacman()
