"""Golden vectors for Game.update_players, produced by the REFERENCE's own C++ (build container only).

    python tests/golden/make_golden_helpers.py

oracle/build_ref.py compiles lines 15-66, 146-156, 191-279 of /root/reference/games/game_helpers.cpp (the Boost-free,
torch-free part) into oracle/_ref/libgame_ref.so; this script drives it and writes tests/golden/game_update.npz:

  A  `agents`: the positions the shipped agents drove (fixture race_agents.npz), every car every 2nd step, rows of
     2 floats (the new position);
  B  `lagged`: the reference's real call pattern (games/race.py:394, 423): rows (old_x, old_y, new_x, new_y) of the
     cars that are alive, row stride 4 - update_players reads columns 0, 1;
  C  `jumps`: random jumps forwards, backwards, across walls and behind the start line on random iid-9 tracks
     (exercises the backward walk, "next_seg < 0 -> dead", finishing, re-entering after death);
  K  `predicates`: orientation / segment_intersect on random and degenerate (collinear, touching) point sets.
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class RefGame(object):
    def __init__(self, lib, left, right, P):
        self.lib = lib
        self.left = np.ascontiguousarray(left, dtype=np.float32)
        self.right = np.ascontiguousarray(right, dtype=np.float32)
        lib.ref_game_create.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(lib.ref_game_create(_p(self.left), _p(self.right), left.shape[0], left.shape[1], P))

    def update(self, idx, rows):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        dead = np.zeros(len(idx), dtype=np.uint8)
        fin = np.zeros(len(idx), dtype=np.uint8)
        self.lib.ref_game_update_players(self.h, _p(idx), _p(rows), len(idx), rows.shape[1], _p(dead), _p(fin))
        return dead, fin

    def close(self):
        self.lib.ref_game_destroy(self.h)


def main():
    path = build_ref.build(force=True)
    if path is None:
        sys.exit('the reference is not here (/root/reference): golden vectors can only be made in the build container')
    lib = ctypes.CDLL(path)
    out = {}
    z = np.load(os.path.join(HERE, 'race_agents.npz'))
    left, right, pos, alive = z['left'], z['right'], z['pos'], z['alive']
    B, P = pos.shape[1], pos.shape[2]
    K = B * P
    # A: every car, every 2nd step, rows of 2
    g = RefGame(lib, left, right, P)
    steps = list(range(1, pos.shape[0], 2))
    A_rows = np.stack([pos[s].reshape(K, 2) for s in steps]).astype(np.float32)
    A_dead, A_fin = zip(*[g.update(np.arange(K), A_rows[i]) for i in range(len(steps))])
    g.close()
    out.update(A_left=left, A_right=right, A_P=np.int32(P), A_rows=A_rows, A_dead=np.stack(A_dead), A_fin=np.stack(A_fin))
    # B: alive cars only, rows (old, new), stride 4
    g = RefGame(lib, left, right, P)
    T = min(pos.shape[0] - 1, 120)
    B_idx = np.full((T, K), -1, dtype=np.int64)
    B_rows = np.zeros((T, K, 4), dtype=np.float32)
    B_dead = np.zeros((T, K), dtype=np.uint8)
    B_fin = np.zeros((T, K), dtype=np.uint8)
    for s in range(T):
        sel = np.nonzero(alive[s].reshape(K))[0]
        rows = np.concatenate((pos[s].reshape(K, 2), pos[s + 1].reshape(K, 2)), axis=1)[sel]
        d, f = g.update(sel, rows)
        B_idx[s, :len(sel)] = sel
        B_rows[s, :len(sel)] = rows
        B_dead[s, :len(sel)] = d
        B_fin[s, :len(sel)] = f
    g.close()
    out.update(B_idx=B_idx, B_rows=B_rows, B_dead=B_dead, B_fin=B_fin)
    # C: random jumps on iid-9 tracks
    zi = np.load(os.path.join(HERE, 'race_iid9.npz'))
    cl, cr, cc = zi['left'][:16], zi['right'][:16], zi['centre'][:16]
    rng = np.random.default_rng(77)
    P2, S = 3, 60
    K2 = cl.shape[0] * P2
    g = RefGame(lib, cl, cr, P2)
    C_rows = np.zeros((S, K2, 2), dtype=np.float32)
    C_dead = np.zeros((S, K2), dtype=np.uint8)
    C_fin = np.zeros((S, K2), dtype=np.uint8)
    prog = np.zeros(K2, dtype=np.int64)
    for s in range(S):
        kind = rng.integers(0, 10, K2)
        step = np.where(kind < 6, rng.integers(0, 4, K2), np.where(kind < 8, -rng.integers(1, 4, K2), rng.integers(4, 40, K2)))
        prog = np.clip(prog + step, -1, cc.shape[1] - 1)
        trk = np.arange(K2) // P2
        base = np.where(prog[:, None] >= 0, cc[trk, np.maximum(prog, 0)], np.array([[0., -0.3]]))
        lateral = np.where(kind[:, None] == 9, 3.0, 0.25) * rng.standard_normal((K2, 2))
        C_rows[s] = (base + lateral).astype(np.float32)
        C_dead[s], C_fin[s] = g.update(np.arange(K2), C_rows[s])
    g.close()
    out.update(C_left=cl, C_right=cr, C_P=np.int32(P2), C_rows=C_rows, C_dead=C_dead, C_fin=C_fin)
    # K: predicates
    pts = rng.standard_normal((400, 4, 2)).astype(np.float32)
    pts[100:200] = np.round(pts[100:200] * 2) / 2                       # lattice points: collinear / touching cases
    pts[200:250, 2] = pts[200:250, 0] + (pts[200:250, 1] - pts[200:250, 0]) * 0.5      # p on the segment a-b
    pts[250:300, 3] = pts[250:300, 1]                                                  # shared end point
    lib.ref_orientation.argtypes = [ctypes.c_float] * 6
    K_orient = np.array([lib.ref_orientation(*map(float, (q[0, 0], q[0, 1], q[1, 0], q[1, 1], q[2, 0], q[2, 1]))) for q in pts], dtype=np.int32)
    K_cross = np.array([lib.ref_segment_intersect(_p(np.ascontiguousarray(q[0])), _p(np.ascontiguousarray(q[1])),
                                                  _p(np.ascontiguousarray(q[2])), _p(np.ascontiguousarray(q[3]))) for q in pts], dtype=np.uint8)
    out.update(K_pts=pts, K_orient=K_orient, K_cross=K_cross)
    np.savez_compressed(os.path.join(HERE, 'game_update.npz'), **out)
    print('A: %d steps x %d cars, dead %d fin %d' % (len(steps), K, int(np.sum(A_dead)), int(np.sum(A_fin))))
    print('B: %d steps, dead %d fin %d' % (T, int(B_dead.sum()), int(B_fin.sum())))
    print('C: %d steps x %d cars, dead %d fin %d' % (S, K2, int(C_dead.sum()), int(C_fin.sum())))
    print('K: orient %s, cross %d of %d' % (np.bincount(K_orient + 1).tolist(), int(K_cross.sum()), len(K_cross)))


if __name__ == '__main__':
    main()
