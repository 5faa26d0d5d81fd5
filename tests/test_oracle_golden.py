"""The CPU oracle (oracle/race_oracle.py) must reproduce the reference-generated fixtures bit-for-bit.

The fixtures under tests/golden/ were produced by the real reference (tests/golden/make_golden.py);
this is what pins the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import race_oracle as ro
from tests.helpers import RACE_CASES, eq, load_case, t, GOLDEN
import os


def make_oracle(case):
    cars = [ro.Car(*c) for c in case['cars'].tolist()]
    return ro.RaceOracle(timeout=float(case['timeout']), cars=cars,
                         framerate=float(case['framerate']))


@pytest.mark.parametrize('name', RACE_CASES)
def test_oracle_replays_reference(name):
    c = load_case(name)
    env = make_oracle(c)
    states, any_valid = env.reset(t(c['tracks']))
    assert any_valid == bool(c['any_valid'])
    assert env.steps_limit == int(c['steps_limit'])
    for k, ref in (('centre', env.centre), ('left', env.left), ('right', env.right)):
        assert eq(ref, t(c[k])), 'geometry %s differs' % k
    assert eq(env.valid, t(c['valid']))
    assert eq(states, t(c['states'][0]))
    assert env.finished() == bool(c['finished'][0])
    T = c['actions'].shape[0]
    for s in range(T):
        states, rewards = env.step(t(c['actions'][s]))
        w = int(c['widths'][s + 1])
        assert states.size(-1) == w, 'step %d state width' % s
        assert eq(states, t(c['states'][s + 1][:, :, :w])), 'step %d states' % s
        assert eq(rewards, t(c['rewards'][s])), 'step %d rewards' % s
        for k, v in (('pos', env.pos), ('dir', env.dir), ('speed', env.speed), ('alive', env.alive),
                     ('finishes', env.finishes), ('scores', env.scores)):
            assert eq(v, t(c[k][s + 1])), 'step %d %s' % (s, k)
        assert env.finished() == bool(c['finished'][s + 1])
    assert eq(env.winners(), t(c['winners']))


def test_oracle_predicates_known_answers():
    z = np.load(os.path.join(GOLDEN, 'kat_predicates.npz'))
    segs, probes, rays = t(z['segs']), t(z['probes']), t(z['rays'])
    assert eq(ro.segments_cross(segs, probes), t(z['cross']))
    hit, start_on = ro.crossing_tables(segs, probes)
    assert eq(hit, t(z['hit'])) and eq(start_on, t(z['start_on']))
    assert eq(ro.ray_distances(segs, rays), t(z['dist_lat']))
    assert eq(ro.ray_distances(t(z['fsegs']), t(z['frays'])), t(z['dist_f']))
    lines = t(z['lat_lines'])
    ok = ro.tracks_valid(lines[:, :-1], lines[:, -1:])
    assert eq(ok, t(z['lat_ok']))


def test_oracle_rotation_known_answers():
    z = np.load(os.path.join(GOLDEN, 'kat_rotate.npz'))
    assert eq(ro.rotate(t(z['vecs']), t(z['angles'])), t(z['out']))
