"""ctypes front-end of oracle/race_oracle.c (TEST INFRASTRUCTURE: checker and CPU baseline only).

Mirrors the call shapes of the C ABI in include/glg_b200.h, but on host numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'librace_oracle.so')

MAXP, MAXR = 8, 32


class RaceParams(ctypes.Structure):
    """Layout of glg_race_params (include/glg_b200.h)."""
    _fields_ = [('num_players', ctypes.c_int32), ('num_rays', ctypes.c_int32),
                ('steps_limit', ctypes.c_int32), ('max_distance', ctypes.c_float),
                ('step_penalty', ctypes.c_float), ('drag', ctypes.c_float),
                ('progress_div', ctypes.c_float),
                ('vmax', ctypes.c_float * MAXP),
                ('speed_inc', (ctypes.c_float * 3) * MAXP),
                ('turn_cos', (ctypes.c_float * 3) * MAXP),
                ('turn_sin', (ctypes.c_float * 3) * MAXP),
                ('ray_cos', ctypes.c_float * MAXR), ('ray_sin', ctypes.c_float * MAXR)]


class RaceState(ctypes.Structure):
    _fields_ = [('positions', ctypes.c_void_p), ('directions', ctypes.c_void_p),
                ('speeds', ctypes.c_void_p), ('alive', ctypes.c_void_p),
                ('finishes', ctypes.c_void_p), ('scores', ctypes.c_void_p)]


def build(force=False):
    if force or not os.path.exists(LIB) or \
            os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, 'race_oracle.c')):
        subprocess.check_call(['make', '-s', '-C', HERE])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ro_race_step.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CRace(object):
    """Host-array Race driven by the C oracle; `params` is a filled RaceParams."""

    def __init__(self, params):
        self.pr = params
        self.P = params.num_players
        self.W = params.num_rays + 2

    def reset(self, tracks, sin_table=None, cos_table=None, geometry=None):
        tracks = np.ascontiguousarray(tracks, dtype=np.float32)
        B, L = tracks.shape[0], tracks.shape[1]
        self.B, self.N = B, L + 2
        self.geom = np.zeros((B, 3, self.N, 2), dtype=np.float32)
        if geometry is not None:
            centre, left, right = geometry
            self.geom[:, 0], self.geom[:, 1], self.geom[:, 2] = right, left, centre
        else:
            half = 0 if sin_table is None else (len(sin_table) - 1) // 2
            st = None if sin_table is None else _p(np.ascontiguousarray(sin_table, np.float32))
            ct = None if cos_table is None else _p(np.ascontiguousarray(cos_table, np.float32))
            self._tabs = (sin_table, cos_table)
            lib().ro_track_build(_p(tracks), B, L, st, ct, half, _p(self.geom))
        self.valid = np.zeros(B, dtype=np.uint8)
        lib().ro_track_validate(_p(self.geom), B, self.N, _p(self.valid))
        P = self.P
        self.pos = np.zeros((B, P, 2), np.float32)
        self.dir = np.zeros((B, P, 2), np.float32)
        self.speed = np.zeros((B, P), np.float32)
        self.alive = np.zeros((B, P), np.uint8)
        self.finishes = np.zeros((B, P), np.uint8)
        self.scores = np.zeros((B, P), np.int32)
        self.state = RaceState(*(a.ctypes.data for a in (self.pos, self.dir, self.speed, self.alive,
                                                         self.finishes, self.scores)))
        lib().ro_race_init(self.state, B, P)
        self.steps = 0
        self.n_alive = B * P
        return self.step(np.zeros((P, B), np.int64))[0], bool(self.valid.any())

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int64)
        self.steps += 1
        B, P = self.B, self.P
        if self.n_alive == 0:
            return (np.zeros((P, B, self.W - 1), np.float32),
                    ((1. - self.finishes.astype(np.float32)) * np.float32(self.pr.step_penalty)).T)
        states = np.zeros((P, B, self.W), np.float32)
        rewards = np.zeros((P, B), np.float32)
        self.n_alive = lib().ro_race_step(ctypes.byref(self.pr), _p(self.geom), B, self.N, _p(actions),
                                          _p(self.valid), self.state, self.steps, _p(states), _p(rewards))
        return states, rewards

    def finished(self):
        return self.steps > self.pr.steps_limit or self.n_alive == 0

    def winners(self):
        w = np.zeros(self.B, np.int64)
        lib().ro_race_winners(_p(self.scores), _p(self.finishes), _p(self.valid), self.B, self.P,
                              self.pr.steps_limit, _p(w))
        return w
