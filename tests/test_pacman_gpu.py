"""CUDA Pacman (games/pacman.py replacement) and PytorchWrapper against the reference-generated
fixtures and the numpy oracle; integer grid and rewards are bit-exact."""
import numpy as np
import pytest
import torch

from oracle.pacman_oracle import PacmanOracle
from tests.test_oracle_pacman import PACMAN_CASES, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', PACMAN_CASES)
def test_pacman_replays_reference(name):
    from game_level_gan_b200.games import Pacman
    c = load(name)
    B, H, W, C = c['board'].shape
    P = C - 4
    env = Pacman((H, W), P, batch_size=B)
    obs = env.reset(c['board'].copy())
    assert isinstance(obs, tuple) and len(obs) == P and obs[0].dtype == np.float32
    assert np.array_equal(np.stack(obs), c['first_obs'])
    assert np.array_equal(env.grid, c['grids'][0])
    for t in range(c['actions'].shape[0]):
        obs, rew = env.step(c['actions'][t])
        assert np.array_equal(env.grid, c['grids'][t + 1]), 'grid at step %d' % t
        assert rew[0].dtype == np.float64 and np.array_equal(np.stack(rew), c['rewards'][t]), 'rewards at step %d' % t
    assert np.array_equal(np.stack(obs), c['last_obs'])
    assert np.array_equal(env.players, c['players'])
    assert env.state_shape() == (H, W, 4 + 2 * P) and env.actions == 5


def test_pacman_large_batch_vs_oracle_and_wrapper():
    """2048 random 15x15 boards, 2 players, 25 steps through PytorchWrapper (device API, no host
    round trip) against the numpy oracle."""
    from game_level_gan_b200.games import Pacman, PytorchWrapper
    rng = np.random.default_rng(11)
    B, H, W, P, T = 2048, 15, 15, 2, 25
    fields = rng.choice(4, size=(B, H, W), p=[0.4, 0.5, 0.07, 0.03])
    board = np.zeros((B, H, W, 4 + P), dtype=np.int32)
    np.put_along_axis(board[..., :4], fields[..., None], 1, axis=-1)
    for p, (x, y) in enumerate(((0, 0), (H - 1, W - 1))):
        board[:, x, y, :] = 0
        board[:, x, y, 0] = 1
        board[:, x, y, 4 + p] = 1
    orc = PacmanOracle((H, W), P, 64)
    orc.reset(board[:64].copy())
    env = PytorchWrapper(Pacman((H, W), P, batch_size=B))
    states = env.reset(torch.from_numpy(board).float().cuda())
    assert len(states) == P and states[0].shape == (B, 4 + 2 * P, H, W) and states[0].is_cuda
    total = 0.
    for t in range(T):
        acts = rng.integers(0, 5, size=(B, P)).astype(np.int32)
        states, rewards = env.step(torch.from_numpy(acts).cuda())
        oobs, orew = orc.step(acts[:64])
        assert np.array_equal(rewards[:, :64].cpu().numpy(), np.stack(orew))
        assert np.array_equal(states[1][:64].permute(0, 2, 3, 1).cpu().numpy(), oobs[1])
        total += float(rewards.sum())
    assert np.array_equal(env.grid[:64], orc.grid) and total > 0
    assert env.num_players == P and env.actions == 5          # attribute passthrough
