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
