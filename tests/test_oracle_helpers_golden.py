"""oracle/helpers_oracle.py (the restatement of Game.update_players and its predicates) against
tests/golden/game_update.npz - vectors produced by the REFERENCE's own C++ (game_helpers.cpp:15-66, 146-156,
191-279 compiled where they lie by oracle/build_ref.py, driven by tests/golden/make_golden_helpers.py).
This is what pins that oracle; tests/test_helpers_gpu.py then holds the CUDA entry point to the same vectors."""
import os

import numpy as np
import pytest

from oracle import helpers_oracle as ho
from tests.helpers import GOLDEN


@pytest.fixture(scope='module')
def g():
    z = np.load(os.path.join(GOLDEN, 'game_update.npz'))
    return {k: z[k] for k in z.files}


def replay(game, rows_seq, idx_seq, dead, fin):
    for s in range(len(rows_seq)):
        idx = idx_seq[s]
        d, f = game.update_players(list(idx), rows_seq[s][:len(idx)])
        assert np.array_equal(d, dead[s][:len(idx)]), 'dead, step %d' % s
        assert np.array_equal(f, fin[s][:len(idx)]), 'finished, step %d' % s


def test_agents_positions_every_car(g):
    K = g['A_rows'].shape[1]
    replay(ho.GameOracle(g['A_left'], g['A_right'], int(g['A_P'])), g['A_rows'], [np.arange(K)] * len(g['A_rows']),
           g['A_dead'], g['A_fin'])
    assert g['A_fin'].sum() > 0


def test_reference_call_pattern_alive_cars_rows_of_four(g):
    """games/race.py:394, 423: rows (old, new) of the alive cars; columns 0, 1 are what update_players reads."""
    idx = [r[r >= 0] for r in g['B_idx']]
    replay(ho.GameOracle(g['A_left'], g['A_right'], int(g['A_P'])), g['B_rows'], idx, g['B_dead'], g['B_fin'])


def test_random_jumps_backward_walk_and_cell_minus_one(g):
    """Includes cars that backed out over the start line: the reference's unsigned comparison (int cell index against
    size_t length) then reports them finished and not dead on every later call."""
    K = g['C_rows'].shape[1]
    replay(ho.GameOracle(g['C_left'], g['C_right'], int(g['C_P'])), g['C_rows'], [np.arange(K)] * len(g['C_rows']),
           g['C_dead'], g['C_fin'])
    assert g['C_fin'].sum() > 0 and g['C_dead'].sum() > 0


def test_predicates(g):
    o = [ho.orientation(q[0], q[1], q[2]) for q in g['K_pts']]
    c = [int(ho.segment_intersect(q[0], q[1], q[2], q[3])) for q in g['K_pts']]
    assert np.array_equal(o, g['K_orient']) and np.array_equal(c, g['K_cross'])
    assert (g['K_orient'] == 0).sum() > 20                      # the degenerate cases are in there


def test_restatement_against_the_compiled_reference_live():
    """Where oracle/_ref/libgame_ref.so exists (built by oracle/build_ref.py from the reference's own C++; it travels with
    the working tree), drive it and the restatement side by side on fresh random jumps - not only the committed vectors."""
    import ctypes
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(os.path.dirname(here), 'oracle', '_ref', 'libgame_ref.so')
    if not os.path.exists(path):
        pytest.skip('oracle/_ref/libgame_ref.so not built (python oracle/build_ref.py in the build container)')
    sys.path.insert(0, os.path.join(here, 'golden'))
    from make_golden_helpers import RefGame
    lib = ctypes.CDLL(path)
    z = np.load(os.path.join(GOLDEN, 'race_iid9.npz'))
    left, right, centre = z['left'][4:10], z['right'][4:10], z['centre'][4:10]
    rng = np.random.default_rng(123)
    P, K = 2, 12
    ref = RefGame(lib, left, right, P)
    orc = ho.GameOracle(left, right, P)
    prog = np.zeros(K, dtype=np.int64)
    for s in range(40):
        prog = np.clip(prog + rng.integers(-3, 6, K), -1, centre.shape[1] - 1)
        trk = np.arange(K) // P
        base = np.where(prog[:, None] >= 0, centre[trk, np.maximum(prog, 0)], np.array([[0., -0.3]]))
        rows = (base + 0.3 * rng.standard_normal((K, 2))).astype(np.float32)
        sel = np.nonzero(rng.random(K) < 0.8)[0]
        d, f = ref.update(sel, rows[sel])
        od, of = orc.update_players(list(sel), rows[sel])
        assert np.array_equal(d, od) and np.array_equal(f, of), 'step %d' % s
    ref.close()
