"""Pure-Python oracle for the `game_helpers` entry points (TEST INFRASTRUCTURE - checker only).

Game.update_players is Boost-free in the reference and is restated literally (game_helpers.cpp:191-279,
with orientation / on_segment / segment_intersect of :22-66) in numpy float32 arithmetic - PINNED against
tests/golden/game_update.npz, which the reference's own C++ produced (oracle/build_ref.py compiles those line
ranges of /root/reference/games/game_helpers.cpp where they lie; tests/golden/make_golden_helpers.py drives it).
The Boost.Geometry-backed functions (collision, smallest_distance, is_valid, Game.validate_tracks,
Game.smallest_distance) have no golden vectors upstream and Boost is not available here:
PARITY UNPINNED for those - they are defined as in csrc/glg_helpers.cu and checked here against the
same definitions written independently in float64, away from degeneracies, plus the one hand-derived
known-answer case of games/run_game_helpers.py:10-28.
"""
import math

import numpy as np

f32 = np.float32


def orientation(a, b, r):
    """game_helpers.cpp:22-37"""
    bax, bay = f32(b[0] - a[0]), f32(b[1] - a[1])
    rbx, rby = f32(r[0] - b[0]), f32(r[1] - b[1])
    v = f32(f32(bay * rbx) - f32(bax * rby))
    return int(v > 0) - int(v < 0)


def on_segment(a, b, r):
    """game_helpers.cpp:39-47"""
    return min(a[0], b[0]) <= r[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= r[1] <= max(a[1], b[1])


def segment_intersect(a, b, p, q):
    """game_helpers.cpp:49-66"""
    o1, o2 = orientation(a, b, p), orientation(a, b, q)
    o3, o4 = orientation(p, q, a), orientation(p, q, b)
    if o1 != o2 and o3 != o4:
        return True
    return ((o1 == 0 and on_segment(a, b, p)) or (o2 == 0 and on_segment(a, b, q)) or
            (o3 == 0 and on_segment(p, q, a)) or (o4 == 0 and on_segment(p, q, b)))


class GameOracle(object):
    """game_helpers.cpp:146-174 (Player, Game constructor) and :191-279 (update_players)."""

    def __init__(self, left, right, num_players):
        self.left = np.asarray(left, dtype=f32)
        self.right = np.asarray(right, dtype=f32)
        self.P = num_players
        b = self.left.shape[0]
        self.pos = np.tile(np.array([0.0, 0.1], dtype=f32), (b * num_players, 1))
        self.seg = np.zeros(b * num_players, dtype=np.int64)

    def update_players(self, idx, new_positions):
        dead, fin = [], []
        length = self.left.shape[1] - 1
        for i, pl in enumerate(idx):
            L, R = self.left[pl // self.P], self.right[pl // self.P]
            new = np.asarray(new_positions[i][:2], dtype=f32)
            old = self.pos[pl].copy()
            nxt = int(self.seg[pl])
            alive, done = True, False
            # `next_seg` is an int, `track.length` a size_t: the two comparisons of the forward part are unsigned in the
            # reference (game_helpers.cpp:105, 215, 220, 240) - a car at cell -1 skips the walk and is "finished"
            while 0 <= nxt < length:
                if segment_intersect(L[nxt], L[nxt + 1], old, new) or segment_intersect(R[nxt], R[nxt + 1], old, new):
                    alive = False
                    break
                if orientation(L[nxt + 1], R[nxt + 1], new) > 0:
                    break
                nxt += 1
            if alive and (nxt >= length or nxt < 0):
                done = True
            if nxt == self.seg[pl] and alive and not done:
                while nxt >= 0:
                    if segment_intersect(L[nxt], L[nxt + 1], old, new) or segment_intersect(R[nxt], R[nxt + 1], old, new):
                        alive = False
                        break
                    if orientation(L[nxt], R[nxt], new) < 0:
                        break
                    nxt -= 1
                if nxt < 0:
                    alive = False
            self.pos[pl] = new
            self.seg[pl] = nxt
            dead.append(0 if alive else 1)
            fin.append(1 if done else 0)
        return np.array(dead, dtype=np.uint8), np.array(fin, dtype=np.uint8)

    def line(self, trk):
        """left reversed, then right (game_helpers.cpp:127-138)"""
        return np.concatenate((self.left[trk][::-1], self.right[trk]), axis=0)


def _cross64(ax, ay, bx, by):
    return ax * by - ay * bx


def ray_distance64(line, ray):
    """Nearest common point of the polyline and the segment origin -> origin + 1000*d, in float64,
    for non-degenerate configurations (no collinear overlap)."""
    sx, sy, dx, dy = (float(v) for v in ray)
    fx, fy = sx + 1000.0 * dx, sy + 1000.0 * dy
    best = math.inf
    pts = np.asarray(line, dtype=np.float64)
    for (px, py), (qx, qy) in zip(pts[:-1], pts[1:]):
        den = _cross64(fx - sx, fy - sy, qx - px, qy - py)
        if den == 0.0:
            continue
        t = _cross64(px - sx, py - sy, qx - px, qy - py) / den
        u = _cross64(px - sx, py - sy, fx - sx, fy - sy) / den
        if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
            best = min(best, t * math.hypot(fx - sx, fy - sy))
    return best


def polyline_hits64(line, seg):
    pts = np.asarray(line, dtype=np.float64)
    a = (float(seg[0]), float(seg[1]))
    b = (float(seg[2]), float(seg[3]))
    return any(segment_intersect(p, q, a, b) for p, q in zip(pts[:-1], pts[1:]))


def self_intersects64(line):
    pts = [tuple(float(v) for v in p) for p in line]
    m = len(pts) - 1
    for i in range(m):
        if i + 1 < m:
            a, c, e = pts[i], pts[i + 1], pts[i + 2]
            if (orientation(a, c, e) == 0 and on_segment(a, c, e)) or (orientation(c, e, a) == 0 and on_segment(c, e, a)):
                return True
        for j in range(i + 2, m):
            if segment_intersect(pts[i], pts[i + 1], pts[j], pts[j + 1]):
                return True
    return False
