"""Fuzz: the production kernels (pruned sensors / collision: per-step packed kernel and the fused rollout kernel)
against the literal kernel (GLG_STEP_BRUTE: every ray x every wall, every wall x path, games/race.py:213-308) on
adversarial states - >= 1e8 car-steps by default.

    python tools/fuzz_pruned_vs_brute.py [--car-steps 1e8] [--out profiles/r02_fuzz_pruned_vs_brute.json]

Each round draws a batch of tracks (iid-9 generator levels / arbitrary float arcs and widths / gentle float tracks),
1..8 cars, resets three environments (brute, fast per-step, fast fused) and then, every `T` steps, THROWS the cars to
random states: positions at centre points plus lateral noise - some exactly on wall vertices or on a wall's line, some
far outside the track ("off-origin") -, headings of random angle whose norm has drifted from 1 (0.6 .. 1.5, a few beyond
the pruning's precondition so the in-kernel fallback runs), random speeds.  All three play the same random actions;
every observation, reward and state array of every step must be bit-equal (NaN == NaN).  Mismatches are counted and the
first ones dumped with their distance to the nearest wall line ("near boundary" = within 1e-5).
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from game_level_gan_b200.games import Race, RaceCar  # noqa: E402

CARS = [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.), (70., 5., 30.), (40., 2., 90.), (90., 1., 45.), (55., 3., 70.)]


def same(a, b):
    return (a == b) | (torch.isnan(a) & torch.isnan(b)) if a.dtype.is_floating_point else a == b


def make_tracks(kind, B, L, gen, dev):
    tr = torch.zeros(B, L, 2, device=dev)
    if kind == 'iid9':
        tr[:, :, 0] = torch.linspace(-1., 1., 9, device=dev)[torch.randint(0, 9, (B, L), generator=gen, device=dev)]
    elif kind == 'float':
        tr[:, :, 0] = torch.rand((B, L), generator=gen, device=dev) * 2 - 1
        tr[:, :, 1] = torch.rand((B, L), generator=gen, device=dev)
    else:                                                                    # gentle float arcs, float widths
        tr[:, :, 0] = (torch.rand((B, L), generator=gen, device=dev) * 2 - 1) * 0.3
        tr[:, :, 1] = torch.rand((B, L), generator=gen, device=dev)
    return tr


def throw_cars(env, gen, dev):
    """random car states, written into the environment's own arrays"""
    B, P = env.num_tracks, env.num_players
    centre, left, right = env.segments, env.left_vecs, env.right_vecs        # [B, N, 2]
    N = centre.size(1)
    j = torch.randint(0, N, (B, P), generator=gen, device=dev)
    pick = lambda x: torch.gather(x, 1, j[..., None].expand(-1, -1, 2))
    c, l, r = pick(centre), pick(left), pick(right)
    kind = torch.rand((B, P), generator=gen, device=dev)
    lam = torch.rand((B, P, 1), generator=gen, device=dev)
    pos = l + (r - l) * lam                                                  # somewhere across the track
    pos = torch.where((kind < 0.10)[..., None], l, pos)                      # exactly on a wall vertex
    pos = torch.where(((kind >= 0.10) & (kind < 0.15))[..., None], l + (r - l) * torch.round(lam * 4) / 4, pos)
    jn = (j + 1).clamp(max=N - 1)
    ln = torch.gather(left, 1, jn[..., None].expand(-1, -1, 2))
    pos = torch.where(((kind >= 0.15) & (kind < 0.25))[..., None], l + (ln - l) * (lam * 3 - 1), pos)   # on a wall's LINE
    far = (kind >= 0.25) & (kind < 0.30)
    pos = torch.where(far[..., None], c + (torch.rand((B, P, 2), generator=gen, device=dev) - 0.5) * 60., pos)   # off the track
    very_far = kind >= 0.995
    pos = torch.where(very_far[..., None], pos * 40. + 150., pos)            # beyond the pruning's 200-unit precondition
    ang = torch.rand((B, P), generator=gen, device=dev) * 6.2831853
    norm = 0.6 + 0.9 * torch.rand((B, P), generator=gen, device=dev)
    norm = torch.where(torch.rand((B, P), generator=gen, device=dev) < 0.6, torch.ones_like(norm) + (norm - 1.05) * 1e-3, norm)
    norm = torch.where(torch.rand((B, P), generator=gen, device=dev) < 0.01, norm * 3., norm)    # precondition fails
    dirs = torch.stack((torch.sin(ang), torch.cos(ang)), -1) * norm[..., None]
    spd = torch.rand((B, P), generator=gen, device=dev) * env.cars_max_speed[None, :].to(dev) * 1.2
    return pos, dirs, spd


def set_state(env, pos, dirs, spd):
    env.positions.copy_(pos)
    env.directions.copy_(dirs)
    env.speeds.copy_(spd)
    env.alive.fill_(True)
    env.finishes.fill_(False)
    env.scores.zero_()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--car-steps', type=float, default=1e8)
    ap.add_argument('--tracks', type=int, default=8192)
    ap.add_argument('--steps', type=int, default=6, help='steps between two throws')
    ap.add_argument('--throws', type=int, default=8, help='throws per batch of tracks')
    ap.add_argument('--seed', type=int, default=2024)
    ap.add_argument('--out', default=None)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(args.seed)
    total = alive_steps = values = mism = near = rounds = 0
    by = {}
    examples = []
    t0 = time.time()
    while total < args.car_steps:
        kind = ('iid9', 'float', 'gentle')[rounds % 3]
        P = 1 + (rounds // 3) % 8
        L = (128, 128, 64, 200)[(rounds // 24) % 4]
        B, T = args.tracks, args.steps
        cars = [RaceCar(*c) for c in CARS[:P]]
        tracks = make_tracks(kind, B, L, gen, dev)
        envs = {v: Race(timeout=40., cars=cars, framerate=1. / 20., log_history=False, variant=var, device=dev)
                for v, var in (('brute', 'brute'), ('step', 'fast'), ('fused', 'fast'))}
        for e in envs.values():
            e.reset(tracks)
        ref = envs['brute']
        for throw in range(args.throws):
            pos, dirs, spd = throw_cars(ref, gen, dev)
            for e in envs.values():
                set_state(e, pos, dirs, spd)
            acts = torch.randint(0, 9, (T, P, B), generator=gen, device=dev)
            out_b = [ref.step(acts[s]) for s in range(T)]
            st_b, rw_b = torch.stack([o[0] for o in out_b]), torch.stack([o[1] for o in out_b])
            alive_steps += int(sum((o[0][:, :, :18] > 0).any(-1).sum() for o in out_b))
            out_s = [envs['step'].step(acts[s]) for s in range(T)]
            st_s, rw_s = torch.stack([o[0] for o in out_s]), torch.stack([o[1] for o in out_s])
            st_f, rw_f = envs['fused'].rollout(acts, keep_all=True, mode='fused')
            for name, st, rw, e in (('step', st_s, rw_s, envs['step']), ('fused', st_f, rw_f, envs['fused'])):
                bad = (~same(st, st_b)).any(-1) | ~same(rw, rw_b)                      # [T, P, B]
                for x, y in ((e.positions, ref.positions), (e.directions, ref.directions)):
                    bad[-1] |= (~same(x, y)).any(-1).t()
                for x, y in ((e.speeds, ref.speeds), (e.alive, ref.alive), (e.finishes, ref.finishes), (e.scores, ref.scores)):
                    bad[-1] |= (~same(x, y)).t()
                n = int(bad.sum())
                values += st.numel() + rw.numel()
                if n:
                    mism += n
                    by[(name, kind, P, L)] = by.get((name, kind, P, L), 0) + n
                    for idx in bad.nonzero()[:5].tolist():
                        if len(examples) < 20:
                            s_, p_, b_ = idx
                            examples.append({'kernel': name, 'tracks': kind, 'P': P, 'L': L, 'round': rounds, 'throw': throw,
                                             'step': s_, 'player': p_, 'board': b_,
                                             'got': st[s_, p_, b_].tolist(), 'want': st_b[s_, p_, b_].tolist()})
            total += T * P * B
        rounds += 1
        if rounds % 6 == 0:
            print('round %d: %.3g car-steps, mismatching car-steps %d (%.1f s)' % (rounds, total, mism, time.time() - t0), flush=True)
    res = {'tool': 'tools/fuzz_pruned_vs_brute.py', 'car_steps': total, 'car_steps_with_a_reading': alive_steps,
           'values_compared': values, 'rounds': rounds, 'kernels': ['race_step_packed_kernel (per step)', 'race_rollout_fused_kernel'],
           'against': 'race_step_kernel<BRUTE> (literal loops)', 'mismatching_car_steps': mism, 'near_boundary': near,
           'by_config': {str(k): v for k, v in by.items()}, 'examples': examples, 'seconds': time.time() - t0,
           'inputs': 'iid-9 / float / gentle-float tracks, L in {64, 128, 200}, P = 1..8, cars thrown every %d steps: on wall '
                     'vertices, on wall lines, across the track, up to 30 units off the track, beyond 200 units; heading norm '
                     '0.6..1.5 (1 %% x3: precondition fails); random speeds and actions' % args.steps}
    print(json.dumps(res))
    if args.out:
        with open(args.out, 'w') as f:
            json.dump(res, f, indent=1)
    sys.exit(1 if mism else 0)


if __name__ == '__main__':
    main()
