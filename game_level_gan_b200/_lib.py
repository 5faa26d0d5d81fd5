"""ctypes binding of the C ABI (include/glg_b200.h) - the only path from Python to the kernels.

`lib()` loads game_level_gan_b200/csrc/libglg_b200.so (built by game_level_gan_b200/build.py, i.e.
`__graft_entry__.build()`).  A missing library is a hard error: there is no fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GLG_LIB_PATH') or os.path.join(HERE, 'csrc', 'libglg_b200.so')    # (override: A/B builds, tools/ab.py)

MAX_PLAYERS = 8
MAX_RAYS = 32
ALIVE_SLOTS = 1024
STEP_FAST, STEP_BRUTE, STEP_SCAN, STEP_PACKED = 0, 1, 2, 3
ROLLOUT_STEPWISE, ROLLOUT_CHAINED, ROLLOUT_FUSED = 0, 1, 2
ABI_VERSION = 7


class GlgError(RuntimeError):
    pass


class RaceParams(ctypes.Structure):
    """glg_race_params"""
    _fields_ = [('num_players', ctypes.c_int32), ('num_rays', ctypes.c_int32),
                ('steps_limit', ctypes.c_int32), ('max_distance', ctypes.c_float),
                ('step_penalty', ctypes.c_float), ('drag', ctypes.c_float),
                ('progress_div', ctypes.c_float),
                ('vmax', ctypes.c_float * MAX_PLAYERS),
                ('speed_inc', (ctypes.c_float * 3) * MAX_PLAYERS),
                ('turn_cos', (ctypes.c_float * 3) * MAX_PLAYERS),
                ('turn_sin', (ctypes.c_float * 3) * MAX_PLAYERS),
                ('ray_cos', ctypes.c_float * MAX_RAYS), ('ray_sin', ctypes.c_float * MAX_RAYS)]


class RaceState(ctypes.Structure):
    """glg_race_state"""
    _fields_ = [('positions', ctypes.c_void_p), ('directions', ctypes.c_void_p),
                ('speeds', ctypes.c_void_p), ('alive', ctypes.c_void_p),
                ('finishes', ctypes.c_void_p), ('scores', ctypes.c_void_p)]


# name -> (restype, argtypes); every symbol include/glg_b200.h declares
_vp, _i32, _i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
SYMBOLS = {
    'glg_last_error': (ctypes.c_char_p, []),
    'glg_abi_version': (ctypes.c_int, []),
    'glg_track_build': (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    'glg_track_build_levels': (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    'glg_track_validate': (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp]),
    'glg_track_validate_pairs': (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp]),
    'glg_track_extent': (ctypes.c_int, [_vp, _i32, _i32, _vp, _vp]),
    'glg_race_init': (ctypes.c_int, [RaceState, _i32, _i32, _vp, _vp]),
    'glg_race_step': (ctypes.c_int, [ctypes.POINTER(RaceParams), _vp, _i32, _i32, _vp, _vp, _vp, RaceState,
                                     _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp]),
    'glg_race_rollout': (ctypes.c_int, [ctypes.POINTER(RaceParams), _vp, _i32, _i32, _vp, _i32, _vp, _vp,
                                        RaceState, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    'glg_race_chain_bytes': (_i64, [_i32, _i32]),
    'glg_race_winners': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    'glg_winner_stats': (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp]),
    'glg_collision': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    'glg_smallest_distance': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    'glg_is_valid': (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp]),
    'glg_game_workspace_bytes': (_i64, [_i32, _i32, _i32]),
    'glg_game_create': (ctypes.c_int, [ctypes.POINTER(_vp), _vp, _i64, _vp, _vp, _i32, _i32, _i32, _vp]),
    'glg_game_destroy': (None, [_vp]),
    'glg_game_validate_tracks': (ctypes.c_int, [_vp, _vp, _vp]),
    'glg_game_update_players': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    'glg_game_smallest_distance': (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    'glg_pacman_step': (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    'glg_pacman_observe': (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
}

_lib = None


def lib():
    """The loaded C-ABI library; raises GlgError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GlgError('CUDA library %s is missing - run `python -c "import __graft_entry__ as g; '
                           'g.build()"` (there is no CPU fallback)' % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.glg_abi_version() != ABI_VERSION:
            raise GlgError('libglg_b200.so ABI version mismatch - rebuild')
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().glg_last_error()
        raise GlgError('%s failed (%d): %s' % (what, code, msg.decode() if msg else '?'))


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), 'glg kernels need contiguous tensors'
    return t.data_ptr()


def stream_ptr(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
