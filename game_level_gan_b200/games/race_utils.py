"""Default car parameters, the module-level `race_game` and the hand-made tracks
(mirrors games/race_utils.py:13-56 of the reference)."""
import random

import torch

from .race import Race, RaceCar


class RaceConfig(object):
    max_segments = 128
    cars = [RaceCar(max_speed=60., acceleration=4., angle=40.),
            RaceCar(max_speed=60., acceleration=1., angle=80.)]


# constructing a Race touches no device; the first reset() binds it to the current CUDA device
race_game = Race(timeout=40., framerate=1. / 20., cars=RaceConfig.cars)


def predefined_tracks(device=None):
    """Six hand-made tracks with random offsets, [6, max_segments, 2] (arc in column 0, width 0).

    Draws from Python's `random` in the same order as the reference, so `random.seed(s)` gives the
    same tracks."""
    L = RaceConfig.max_segments
    arcs = torch.zeros(6, L)

    def span(row, start, length, value):
        lo, hi = max(start, 0), min(start + length, L)
        if hi > lo:
            arcs[row, lo:hi] = value

    zig = random.randint(0, 15)                      # alternating 16-segment bends
    for i in range(zig, L, 16):
        span(0, i, 16, 2. * ((i // 16) % 2) - 1.)
    span(1, random.randint(0, 100), 20, 1.)          # one sharp turn
    s = random.randint(0, 70)                        # S bend
    span(2, s, 12, 1.)
    span(2, s + 30, 12, -1.)
    jitter = [random.randint(0, 5) for _ in range(7)]   # wide U turn in 7 pieces
    for off, j in zip((0, 10, 20, 40, 50, 70, 100), jitter):
        span(3, off + j, 5, 1.)
    bump = random.randint(0, 90)                     # small bumpy turn
    span(4, bump, 6, 1.)
    span(4, bump + 6, 12, -1.)
    span(4, bump + 18, 6, 1.)
    back = random.randint(30, 60)                    # immediate turn, counter-turn near the end
    span(5, 0, 20, 1.)
    span(5, L - back, 20, -1.)
    tracks = torch.zeros(6, L, 2)
    tracks[:, :, 0] = arcs
    if device is None and torch.cuda.is_available():
        device = torch.device('cuda')
    return tracks.to(device) if device is not None else tracks
