"""The plain-C oracle (oracle/race_oracle.c) against the reference-generated fixtures.

This pins the scalar fp32 model (three-rounding predicates, fma norm, double cumsum) that the CUDA
kernels implement; the steering / ray tables come from the product's host table builder, so this
also pins game_level_gan_b200/games/_tables.py.
"""
import ctypes

import numpy as np
import pytest

from oracle import c_oracle
from oracle import race_oracle as ro
from game_level_gan_b200.games import _tables
from tests.helpers import RACE_CASES, eq, load_case, nmismatch


def params_for(case):
    cars = [ro.Car(*c) for c in case['cars'].tolist()]
    pr = _tables.race_params(cars, float(case['framerate']), float(case['timeout']), 18, 10.)
    out = c_oracle.RaceParams()
    ctypes.memmove(ctypes.byref(out), ctypes.byref(pr), ctypes.sizeof(out))
    return out


QUANTISED = {'predef', 'iid9', 'loops', 'p1_crash', 'p4_short', 'agents'}


@pytest.mark.parametrize('name', RACE_CASES)
def test_c_oracle_replays_reference(name):
    c = load_case(name)
    env = c_oracle.CRace(params_for(c))
    L = c['tracks'].shape[1]
    st, ct, _ = _tables.heading_tables(L)
    if name in QUANTISED:
        states, any_valid = env.reset(c['tracks'], st.numpy(), ct.numpy())
        # geometry itself must be bit-identical when the table applies
        assert eq(env.geom[:, 2], c['centre']) and eq(env.geom[:, 1], c['left']) and eq(env.geom[:, 0], c['right'])
    else:
        # libm sinf/cosf vs SLEEF: build parity is "within a few ulp"; step parity uses reference geometry
        g_env = c_oracle.CRace(params_for(c))
        g_env.reset(c['tracks'])
        for k, i in (('centre', 2), ('left', 1), ('right', 0)):
            np.testing.assert_allclose(g_env.geom[:, i], c[k], rtol=0, atol=2e-5)
        states, any_valid = env.reset(c['tracks'], geometry=(c['centre'], c['left'], c['right']))
    assert any_valid == bool(c['any_valid'])
    assert eq(np.repeat(env.valid, env.P), c['valid'].astype(np.uint8))
    assert eq(states, c['states'][0])
    for s in range(c['actions'].shape[0]):
        states, rewards = env.step(c['actions'][s])
        w = int(c['widths'][s + 1])
        assert states.shape[-1] == w
        assert nmismatch(states, c['states'][s + 1][:, :, :w]) == 0, 'step %d states' % s
        assert eq(rewards, c['rewards'][s]), 'step %d rewards' % s
        for k, v in (('pos', env.pos), ('dir', env.dir), ('speed', env.speed), ('scores', env.scores)):
            assert eq(v, c[k][s + 1]), 'step %d %s' % (s, k)
        assert eq(env.alive, c['alive'][s + 1].astype(np.uint8))
        assert eq(env.finishes, c['finishes'][s + 1].astype(np.uint8))
        assert env.finished() == bool(c['finished'][s + 1])
    assert eq(env.winners(), c['winners'])
