"""Host-side constant tables for the Race kernels (filled into glg_race_params).

The reference evaluates cos/sin of the steering angle and of the 18 sensor angles with torch on
every step (games/race.py:310-324, 362-363, 462-466).  Those angles take 3 values per car type and
18 fixed values, so they are evaluated here ONCE, on the host, with the very same torch ops - the
kernels then reproduce headings and ray directions bit-for-bit without calling sinf/cosf.
"""
import math

import torch

from .._lib import MAX_PLAYERS, MAX_RAYS, RaceParams

STEER_FLAGS = (0., 1., -1.)     # action_dirs values, games/race.py:62-71
THROTTLE_FLAGS = (0., 1., -3.)  # action_speed values, games/race.py:52-61
# action -> index into the two flag tuples above
ACTION_STEER = (0, 0, 0, 1, 1, 1, 2, 2, 2)
ACTION_THROTTLE = (0, 1, 2, 0, 1, 2, 0, 1, 2)


def _cos_sin(angles):
    """cos/sin exactly as the reference's 2x2 rotation builder evaluates them (strided in-place
    ops on a [n,2,2] buffer, games/race.py:318-322)."""
    m = angles.view(-1, 1).repeat(1, 4).view(-1, 2, 2)
    m[:, 0, 0].cos_()
    m[:, 1, 0].sin_()
    return m[:, 0, 0].clone(), m[:, 1, 0].clone()


def race_params(cars, framerate, timeout, observation_size, max_distance, step_penalty=-0.01,
                drag=0.05):
    P = len(cars)
    if not 1 <= P <= MAX_PLAYERS:
        raise ValueError('number of cars must be in 1..%d' % MAX_PLAYERS)
    if not 1 <= observation_size <= MAX_RAYS:
        raise ValueError('observation_size must be in 1..%d' % MAX_RAYS)
    cpu = torch.device('cpu')
    vmax = torch.tensor([c.max_speed for c in cars], dtype=torch.float32, device=cpu)
    accel = torch.tensor([c.acceleration for c in cars], dtype=torch.float32, device=cpu)
    angle = torch.tensor([c.angle for c in cars], dtype=torch.float32, device=cpu)
    steer = torch.tensor(STEER_FLAGS, device=cpu)
    thr = torch.tensor(THROTTLE_FLAGS, device=cpu)
    pr = RaceParams()
    pr.num_players = P
    pr.num_rays = observation_size
    pr.steps_limit = int(timeout // framerate)          # games/race.py:47
    pr.max_distance = max_distance
    pr.step_penalty = step_penalty
    pr.drag = drag
    pr.progress_div = 3.0                               # games/race.py:350,376: bounds.size(2) - 1
    for p in range(P):
        pr.vmax[p] = float(vmax[p])
        ang = framerate * steer * angle[p].repeat(3)    # games/race.py:363
        c, s = _cos_sin(ang)
        inc = framerate * thr * accel[p].repeat(3)      # games/race.py:367
        for f in range(3):
            pr.turn_cos[p][f] = float(c[f])
            pr.turn_sin[p][f] = float(s[f])
            pr.speed_inc[p][f] = float(inc[f])
    rays = torch.linspace(-math.pi, math.pi * (1. - 2. / observation_size), observation_size,
                          device=cpu)                   # games/race.py:462
    c, s = _cos_sin(rays)
    for i in range(observation_size):
        pr.ray_cos[i] = float(c[i])
        pr.ray_sin[i] = float(s[i])
    return pr


def heading_tables(max_segments):
    """sin/cos of fl32(rad 8 deg) * (n/4) for every integer n a track of `max_segments` arcs in
    0.25 steps can reach (games/race.py:140-142).  Returns (sin, cos, half) with index n + half."""
    half = 4 * max_segments + 8
    n = torch.arange(-half, half + 1, dtype=torch.float32) / 4.
    heading = math.radians(8.) * n
    pad = (-heading.numel()) % 64 + 64                  # keep every entry in the vectorised part
    h = torch.cat((heading, torch.zeros(pad)))
    return torch.sin(h)[:heading.numel()].clone(), torch.cos(h)[:heading.numel()].clone(), half
