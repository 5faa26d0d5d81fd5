"""`game_helpers` facade (free functions + Game) on the GPU against the known-answer case of the
reference's run_game_helpers.py, golden vectors of Game.update_players produced by the reference's own C++
(tests/golden/game_update.npz), the literal Python restatement of it, and float64
restatements of the Boost-backed semantics (parity unpinned upstream, see oracle/helpers_oracle.py)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import helpers_oracle as ho
from tests.helpers import eq, load_case, t

pytestmark = pytest.mark.gpu


def corridor():
    """games/run_game_helpers.py:10-16: two straight tracks, walls x = -1 / +1, y = 0..3."""
    left = torch.zeros(2, 4, 2)
    left[:, :, 0] = -1.
    left[:, :, 1] = torch.linspace(0., 3., 4)[None, :].repeat(2, 1)
    right = left.clone()
    right[:, :, 0] = 1.
    return left, right


def test_run_game_helpers_known_answers():
    """Expected values derived by hand from the reference code."""
    from game_level_gan_b200.games import game_helpers as gh
    left, right = corridor()
    gg = gh.Game(left, right, 2)                       # CPU tensors in -> CPU tensors out
    valid = gg.validate_tracks()
    assert valid.device.type == 'cpu' and valid.dtype == torch.uint8 and valid.tolist() == [1, 1]
    dirs = torch.tensor([[[0., 1., -1., 0.], [0., 1., 0., 1.], [0., 1., 1., 0.], [0., 1., 0., -1.]]])
    d = gg.smallest_distance(torch.tensor([0]), dirs)
    assert d.shape == (1, 4) and d[0, 0] == 1. and math.isinf(d[0, 1]) and d[0, 2] == 1. and d[0, 3] == 1.
    dead, fin = gg.update_players(torch.tensor([0]), torch.tensor([[0., 10.]]))
    assert dead.tolist() == [0] and fin.tolist() == [1]
    # a second player of the same track drives through the left wall
    dead, fin = gg.update_players(torch.tensor([1]), torch.tensor([[-2., 0.5]]))
    assert dead.tolist() == [1] and fin.tolist() == [0]


def test_game_update_players_matches_literal_restatement():
    """Feed the positions the shipped agents drove (fixture `agents`) through Game.update_players and the
    Python restatement of game_helpers.cpp:191-279; dead / finished flags must agree at every step."""
    from game_level_gan_b200.games import game_helpers as gh
    c = load_case('agents')
    left, right = t(c['left']).cuda(), t(c['right']).cuda()
    B, P = left.size(0), 2
    game = gh.Game(left, right, P)
    orc = ho.GameOracle(c['left'], c['right'], P)
    idx = torch.arange(B * P)
    seen_dead = seen_fin = 0
    for s in range(1, c['pos'].shape[0], 3):
        newp = t(c['pos'][s]).reshape(B * P, 2)
        paths = torch.cat((newp, newp), dim=1)          # [k,4] rows: only columns 0,1 are read
        dead, fin = game.update_players(idx.cuda(), paths.cuda())
        od, of = orc.update_players(idx.tolist(), newp.numpy())
        assert eq(dead, od) and eq(fin, of), 'step %d' % s
        seen_dead += int(od.sum())
        seen_fin += int(of.sum())
    assert seen_fin > 0


@pytest.mark.parametrize('case', ['A', 'B', 'C'])
def test_game_update_players_matches_reference_cpp(case):
    """glg_game_update_players against flags the REFERENCE's C++ produced (game_helpers.cpp:191-279 compiled by
    oracle/build_ref.py): A = the agents' positions, every car; B = the reference's own call pattern (alive cars only,
    rows (old, new) of stride 4, games/race.py:394); C = random jumps incl. cars behind the start line (cell -1:
    the unsigned comparison quirk)."""
    from game_level_gan_b200.games import game_helpers as gh
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, 'game_update.npz'))
    src = 'C' if case == 'C' else 'A'
    left, right, P = t(z[src + '_left']).cuda(), t(z[src + '_right']).cuda(), int(z[src + '_P'])
    game = gh.Game(left, right, P)
    rows, dead, fin = z[case + '_rows'], z[case + '_dead'], z[case + '_fin']
    K = rows.shape[1]
    for s in range(rows.shape[0]):
        idx = z['B_idx'][s][z['B_idx'][s] >= 0] if case == 'B' else np.arange(K)
        n = len(idx)
        if n == 0:
            continue
        d, f = game.update_players(t(idx.astype(np.int64)).cuda(), t(rows[s][:n]).cuda())
        assert eq(d, dead[s][:n]) and eq(f, fin[s][:n]), 'case %s step %d' % (case, s)
    assert case == 'B' or fin.sum() > 0
    assert case != 'C' or dead.sum() > 0


def test_stateless_helpers_against_float64_definitions():
    from game_level_gan_b200.games import game_helpers as gh
    c = load_case('iid9')
    g = torch.Generator().manual_seed(3)
    line = torch.cat((t(c['right']).flip(1), t(c['left'])), dim=1)[:6]       # games/race.py:175
    b, s = line.shape[:2]
    centre = t(c['centre'])[:6]
    j = torch.randint(5, 120, (b, 5), generator=g)
    origin = torch.gather(centre, 1, j[..., None].expand(-1, -1, 2)) + torch.randn((b, 5, 2), generator=g) * 0.1
    ang = torch.rand((b, 5), generator=g) * 2 * math.pi
    rays = torch.cat((origin, torch.sin(ang)[..., None], torch.cos(ang)[..., None]), dim=-1)
    out = torch.empty(b, 5)
    gh.smallest_distance(line, rays, out)                                    # CPU in / CPU out
    for i in range(b):
        for k in range(5):
            ref = ho.ray_distance64(line[i].numpy(), rays[i, k].numpy())
            assert (math.isinf(ref) and math.isinf(out[i, k])) or abs(out[i, k].item() - ref) <= 1e-3 * max(1., ref)
    # collision: short probes around the walls
    probes = torch.cat((origin, origin + torch.randn((b, 5, 2), generator=g) * 0.6), dim=-1)
    hit = torch.empty(b, 5, dtype=torch.uint8)
    gh.collision(line, probes, hit)
    ref = [[int(ho.polyline_hits64(line[i].numpy(), probes[i, k].numpy())) for k in range(5)] for i in range(b)]
    assert hit.tolist() == ref and 0 < sum(map(sum, ref)) < b * 5
    # CUDA tensors are used in place
    hit_gpu = torch.zeros(b, 5, dtype=torch.uint8, device='cuda')
    gh.collision(line.cuda(), probes.cuda(), hit_gpu)
    assert hit_gpu.cpu().tolist() == ref


def test_is_valid_and_game_validate():
    from game_level_gan_b200.games import game_helpers as gh
    c = load_case('loops')
    left, right = t(c['left']), t(c['right'])
    line = torch.cat((right.flip(1), left), dim=1)
    out = torch.empty(line.size(0), dtype=torch.uint8)
    gh.is_valid(line, out)
    ref = [0 if ho.self_intersects64(line[i].numpy()) else 1 for i in range(line.size(0))]
    assert out.tolist() == ref
    # away from degeneracies this agrees with the torch path's validity (proper crossings only)
    assert out.tolist() == c['valid'].reshape(-1, 2)[:, 0].astype(int).tolist()
    game = gh.Game(left.cuda(), right.cuda(), 2)
    v = game.validate_tracks()
    assert v.is_cuda and v.cpu().tolist() == ref


def test_argument_errors():
    from game_level_gan_b200.games import game_helpers as gh
    with pytest.raises(RuntimeError):
        gh.is_valid(torch.zeros(2, 4, 2, dtype=torch.float64), torch.empty(2, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        gh.collision(torch.zeros(2, 4, 2), torch.zeros(2, 1, 4), torch.empty(2, 1))
    with pytest.raises(RuntimeError):
        gh.Game(torch.zeros(2, 4, 2), torch.zeros(2, 5, 2), 2)
