"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/glg_b200.h declares; host tables; behaviour without a CUDA device.  No compute calls."""
import ctypes
import os
import random
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'glg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(glg_[a-z_0-9]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as entry
    path = entry.build()
    handle = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), 'missing export %s' % n
    from game_level_gan_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names, 'ctypes table out of sync with the header'
    assert handle.glg_abi_version() == _lib.ABI_VERSION


def test_params_struct_matches_header_size():
    from game_level_gan_b200 import _lib
    # 3 int32 + 4 float + vmax[8] + 3 * [8][3] + 2 * [32]
    assert ctypes.sizeof(_lib.RaceParams) == 4 * (3 + 4 + 8 + 3 * 24 + 64)


def test_argument_validation_without_device():
    """Bad extents are rejected before any CUDA call."""
    from game_level_gan_b200 import _lib
    lib = _lib.lib()
    assert lib.glg_track_build(None, 4, 0, None, None, 0, None, None) == -1
    assert b'glg_track_build' in lib.glg_last_error()
    assert lib.glg_track_validate(None, -1, 130, None, None) == -1
    pr = _lib.RaceParams()
    pr.num_players = 99
    st = _lib.RaceState()
    assert lib.glg_race_step(ctypes.byref(pr), None, 1, 130, None, None, None, st, 1, None, None, None, 1, None, None, 0, 0, None) == -1


def test_host_tables_match_reference_constants():
    from game_level_gan_b200.games import RaceConfig, _tables
    pr = _tables.race_params(RaceConfig.cars, 1. / 20., 40., 18, 10.)
    assert pr.steps_limit == 799 and pr.num_players == 2 and pr.num_rays == 18
    assert abs(pr.vmax[0] - 60. * 100. / 3600.) < 1e-6
    assert pr.turn_cos[0][0] == 1.0 and pr.turn_sin[0][0] == 0.0
    assert pr.turn_sin[0][1] > 0 > pr.turn_sin[0][2] and pr.speed_inc[1][2] < 0 < pr.speed_inc[1][1]
    assert abs(pr.ray_cos[0] + 1.0) < 1e-6 and abs(pr.ray_cos[9] - 1.0) < 1e-6
    s, c, half = _tables.heading_tables(128)
    assert s.numel() == 2 * half + 1 and float(s[half]) == 0.0 and float(c[half]) == 1.0


def test_predefined_tracks_are_seed_reproducible():
    from game_level_gan_b200.games import predefined_tracks
    random.seed(0)
    a = predefined_tracks(device='cpu')
    random.seed(0)
    b = predefined_tracks(device='cpu')
    assert a.shape == (6, 128, 2) and torch.equal(a, b)
    assert set(a[:, :, 0].unique().tolist()) <= {-1., 0., 1.} and float(a[:, :, 1].abs().sum()) == 0.


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-device error path')
def test_no_cpu_fallback():
    from game_level_gan_b200._lib import GlgError
    from game_level_gan_b200.games import Race, RaceConfig
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20.)
    with pytest.raises(GlgError):
        env.reset(torch.zeros(2, 128, 2))
    with pytest.raises(GlgError):
        Race(timeout=40., cars=RaceConfig.cars, device=torch.device('cpu'))
