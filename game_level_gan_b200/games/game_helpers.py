"""Facade of the reference's pybind module `game_helpers` (games/game_helpers.cpp:454-464) over the
sm_100a kernels: the three free functions and class `Game`, same names, argument order, dtypes and
out-parameter conventions.  CUDA tensors are used in place; CPU tensors (what the reference's callers
pass, games/race.py:194-205, 398-401, 474-477) are copied up, and results copied back into the
caller's `out` tensors.  Wrong dtypes / ranks raise RuntimeError like the ATen accessors do.

The Boost.Geometry-backed parts of the reference (intersects, intersection + distance) have no pinned
golden vectors ("parity unpinned"); their semantics here are stated in
csrc/glg_helpers.cu.  Game.update_players is Boost-free upstream and is reproduced literally.
"""
import ctypes

import torch

from .. import _lib
from .._lib import GlgError, check, ptr


def _device():
    if not torch.cuda.is_available():
        raise GlgError('no CUDA device available; game_level_gan_b200 has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _arg(x, dtype, ndim, name):
    if not isinstance(x, torch.Tensor) or x.dtype != dtype or x.dim() != ndim:
        raise RuntimeError('%s: expected a %d-d tensor of %s' % (name, ndim, dtype))
    dev = x.device if x.is_cuda else _device()
    return x.detach().to(dev).contiguous()


def _store(out, result):
    if out.data_ptr() != result.data_ptr():
        out.copy_(result.view(out.shape))


def collision(tracks, segments, output):
    """tracks [b,s,2] f32, segments [b,p,4] f32 -> output [b,p] u8 (in place).  game_helpers.cpp:335-371."""
    t = _arg(tracks, torch.float32, 3, 'tracks')
    sg = _arg(segments, torch.float32, 3, 'segments').to(t.device)
    if not isinstance(output, torch.Tensor) or output.dtype != torch.uint8 or output.dim() != 2:
        raise RuntimeError('output: expected a 2-d uint8 tensor')
    o = output if (output.is_cuda and output.is_contiguous()) else torch.empty(output.shape, dtype=torch.uint8, device=t.device)
    check(_lib.lib().glg_collision(ptr(t), ptr(sg), ptr(o), t.size(0), t.size(1), sg.size(1),
                                   _lib.stream_ptr(t.device)), 'glg_collision')
    _store(output, o)


def smallest_distance(tracks, directions, output):
    """tracks [b,s,2], directions [b,d,(x,y,dx,dy)] -> output [b,d] f32 (in place).  game_helpers.cpp:373-418."""
    t = _arg(tracks, torch.float32, 3, 'tracks')
    dr = _arg(directions, torch.float32, 3, 'directions').to(t.device)
    if not isinstance(output, torch.Tensor) or output.dtype != torch.float32 or output.dim() != 2:
        raise RuntimeError('output: expected a 2-d float32 tensor')
    o = output if (output.is_cuda and output.is_contiguous()) else torch.empty(output.shape, dtype=torch.float32, device=t.device)
    check(_lib.lib().glg_smallest_distance(ptr(t), ptr(dr), ptr(o), t.size(0), t.size(1), dr.size(1),
                                           _lib.stream_ptr(t.device)), 'glg_smallest_distance')
    _store(output, o)


def is_valid(tracks, output):
    """tracks [b,s,2] -> output [b] u8 (in place): 1 where the polyline does not intersect itself.
    game_helpers.cpp:421-450."""
    t = _arg(tracks, torch.float32, 3, 'tracks')
    if not isinstance(output, torch.Tensor) or output.dtype != torch.uint8 or output.dim() != 1:
        raise RuntimeError('output: expected a 1-d uint8 tensor')
    o = output if (output.is_cuda and output.is_contiguous()) else torch.empty(output.shape, dtype=torch.uint8, device=t.device)
    check(_lib.lib().glg_is_valid(ptr(t), ptr(o), t.size(0), t.size(1), _lib.stream_ptr(t.device)), 'glg_is_valid')
    _store(output, o)


class Game(object):
    """Stateful per-player cell tracking, game_helpers.cpp:158-327.  Results are returned on the device
    the boundaries were given on (CPU in, CPU out - like the reference)."""

    def __init__(self, left, right, num_players):
        l = _arg(left, torch.float32, 3, 'left')
        r = _arg(right, torch.float32, 3, 'right').to(l.device)
        if l.shape != r.shape or l.size(2) != 2:
            raise RuntimeError('left/right: expected matching [b,s,2] tensors')
        self._host_io = not left.is_cuda
        self._dev = l.device
        self.b, self.s, self.num_players = l.size(0), l.size(1), int(num_players)
        lib = _lib.lib()
        need = lib.glg_game_workspace_bytes(self.b, self.s, self.num_players)
        if need < 0:
            raise RuntimeError('Game: bad extents')
        self._workspace = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=self._dev)
        handle = ctypes.c_void_p()
        check(lib.glg_game_create(ctypes.byref(handle), ptr(self._workspace), int(need), ptr(l), ptr(r),
                                  self.b, self.s, self.num_players, _lib.stream_ptr(self._dev)), 'glg_game_create')
        self._handle = handle

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h:
            try:
                _lib.lib().glg_game_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass
            self._handle = None

    def _out(self, x):
        return x.cpu() if self._host_io else x

    def validate_tracks(self):
        """-> u8 [b].  game_helpers.cpp:176-189."""
        out = torch.empty((self.b,), dtype=torch.uint8, device=self._dev)
        check(_lib.lib().glg_game_validate_tracks(self._handle, ptr(out), _lib.stream_ptr(self._dev)),
              'glg_game_validate_tracks')
        return self._out(out)

    def update_players(self, idx, new_positions):
        """idx [k] i64, new_positions [k,>=2] f32 -> (dead u8 [k], finished u8 [k]).  game_helpers.cpp:191-279."""
        i = _arg(idx, torch.int64, 1, 'idx').to(self._dev)
        p = _arg(new_positions, torch.float32, 2, 'new_positions').to(self._dev)
        k = i.size(0)
        dead = torch.empty((k,), dtype=torch.uint8, device=self._dev)
        fin = torch.empty((k,), dtype=torch.uint8, device=self._dev)
        check(_lib.lib().glg_game_update_players(self._handle, ptr(i), ptr(p), k, p.size(1), ptr(dead), ptr(fin),
                                                 _lib.stream_ptr(self._dev)), 'glg_game_update_players')
        return self._out(dead), self._out(fin)

    def smallest_distance(self, idx, directions):
        """idx [k] i64, directions [k,d,4] f32 -> f32 [k,d].  game_helpers.cpp:281-322."""
        i = _arg(idx, torch.int64, 1, 'idx').to(self._dev)
        d = _arg(directions, torch.float32, 3, 'directions').to(self._dev)
        out = torch.empty((i.size(0), d.size(1)), dtype=torch.float32, device=self._dev)
        check(_lib.lib().glg_game_smallest_distance(self._handle, ptr(i), ptr(d), i.size(0), d.size(1), ptr(out),
                                                    _lib.stream_ptr(self._dev)), 'glg_game_smallest_distance')
        return self._out(out)
