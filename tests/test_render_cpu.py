"""Drawing helpers against images made by the REFERENCE's own drawing code
(tests/golden/render.npz, tests/golden/make_golden_render.py).  `tracks_images` / `prettier_tracks` only read the
track records, so they are checked here on the CPU with a stand-in for the environment that holds the reference's
geometry of the `iid9` fixture; the episode frames need a live environment (tests/test_render_gpu.py)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, load_case

cv2 = pytest.importorskip('cv2')


class Boards(object):
    """What race_render reads from a Race: `_geom` [B, 3, N, 2] = right boundary reversed, left boundary, centre."""

    def __init__(self, case, n):
        right = torch.from_numpy(case['right'][:n]).flip(1)
        self._geom = torch.stack((right, torch.from_numpy(case['left'][:n]), torch.from_numpy(case['centre'][:n])), 1)
        self.num_players = 2
        self.device = torch.device('cpu')


@pytest.fixture(scope='module')
def golden():
    z = np.load(os.path.join(GOLDEN, 'render.npz'))
    return {k: z[k] for k in z.files}


def test_tracks_images_equal_the_reference(golden):
    from game_level_gan_b200.games import race_render
    imgs = race_render.tracks_images(Boards(load_case('iid9'), 3), top_n=3)
    assert imgs.shape == (3, 256, 256, 3) and imgs.dtype == np.uint8
    assert np.array_equal(imgs, golden['tracks_images'])


def test_prettier_tracks_equal_the_reference(golden):
    from game_level_gan_b200.games import race_render
    imgs = race_render.prettier_tracks(Boards(load_case('iid9'), 2), top_n=2, size=320, pad=0.05)
    assert imgs.shape == (2, 320, 320, 4) and imgs.dtype == np.uint8
    assert np.array_equal(imgs, golden['prettier'])


def test_pixel_transform_follows_the_reference_formula():
    """games/race.py:667-671 evaluated one Python scalar at a time, as the reference does."""
    from game_level_gan_b200.games import race_render
    c = load_case('iid9')
    env = Boards(c, 1)
    walls, finish = race_render.board_segments(env, 0)
    assert walls.shape == (2 * 128 + 3, 4) and finish.shape == (4,)
    assert np.array_equal(walls[0, :2], c['right'][0, 0]) and np.array_equal(walls[-1], np.concatenate((c['left'][0, 0], c['right'][0, 0])))
    assert np.array_equal(finish, np.concatenate((c['left'][0, -1], c['right'][0, -1])))
    b = torch.from_numpy(walls.astype(np.float32))
    mins, maxs = b.view(-1, 2).min(0).values, b.view(-1, 2).max(0).values
    longer = torch.max(maxs - mins).item()
    shiftx, shifty = (0.5 * (1. - (maxs - mins) / longer)).tolist()
    minx, miny = mins.tolist()
    fit = race_render.Fit(walls, 256, 0.05)
    px = fit.pixel(walls.reshape(-1, 2))
    for k, (x, y) in enumerate(b.view(-1, 2)):
        assert px[k, 0] == int(0.05 * 256 + 0.9 * 256 * ((x - minx) / longer + shiftx))
        assert px[k, 1] == 256 - int(0.05 * 256 + 0.9 * 256 * ((y - miny) / longer + shifty))


def test_prettier_tracks_svg_with_a_stand_in_svgwrite(monkeypatch):
    """`svgwrite` is not installed here (the reference imports it inside the method too): a stand-in module records the
    polygons - 129 tarmac quads in two alternating greys (bands of 5) and the 2 x 7 chequered finish strip per board,
    all inside the padded canvas."""
    import sys
    import types
    from game_level_gan_b200.games import race_render

    class Drawing(object):
        def __init__(self, **kw):
            self.kw, self.items = kw, []

        def polygon(self, points, fill):
            return ('polygon', points, fill)

        def add(self, item):
            self.items.append(item)

    fake = types.ModuleType('svgwrite')
    fake.Drawing = Drawing
    fake.rgb = lambda r, g, b: 'rgb(%d,%d,%d)' % (r, g, b)
    monkeypatch.setitem(sys.modules, 'svgwrite', fake)
    imgs = race_render.prettier_tracks_svg(Boards(load_case('iid9'), 2), top_n=2, size=512, pad=0.05)
    assert len(imgs) == 2 and imgs[0].kw == {'shape_rendering': 'crispEdges'}
    for d in imgs:
        assert len(d.items) == 129 + 14
        quads, cells = d.items[:129], d.items[129:]
        assert [q[2] for q in quads[:11]] == ['rgb(70,70,70)'] * 5 + ['rgb(50,50,50)'] * 5 + ['rgb(70,70,70)']
        assert sorted(set(c[2] for c in cells)) == ['rgb(20,20,20)', 'rgb(210,210,210)']
        pts = np.array([p for it in d.items for p in it[1]])
        assert pts.shape[1] == 2 and pts.min() > -0.03 * 512 and pts.max() < 1.03 * 512
