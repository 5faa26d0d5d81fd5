"""Batched grid Pacman on the GPU (replaces games/pacman.py:12-141).

Same constructor, `reset` / `step` / `copy_board` / `from_str` / `state_shape` API as the reference.
The reference keeps everything in numpy; here the grid lives in HBM and `step` is two kernels
(csrc/glg_pacman.cu).  numpy in -> numpy out (drop-in for the reference's own callers);
`reset_device` / `step_device` are the same calls on CUDA tensors with no host round trip, which is
what `PytorchWrapper` uses.
"""
import numpy as np
import torch

from .. import _lib
from .._lib import GlgError, check, ptr
from .environment import MultiEnvironment


class Pacman(MultiEnvironment):
    cuda_native = True

    def __init__(self, size, num_players, batch_size=32, device=None):
        self.size = tuple(size)
        self.fields = 4
        self.depth = self.fields + num_players * 2
        self.grid_depth = self.fields + num_players
        self.num_players = num_players
        self.batch_size = batch_size
        self._device = torch.device(device) if device is not None else None
        self._grid = None          # [B,H,W,4+P] int32 on the device
        self._players = None       # [B*P,4] int32 (board, x, y, player) in np.where order
        self.moves = np.zeros((5, 4), dtype=np.int32)
        self.moves[:, 1:3] = np.array([[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1]], dtype=np.int32)

    @property
    def device(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise GlgError('no CUDA device available; game_level_gan_b200 has no CPU fallback')
            self._device = torch.device('cuda', torch.cuda.current_device())
        return self._device

    def state_shape(self):
        return self.size + (self.depth,)

    def players_layer_shape(self):
        return self.size + (self.num_players,)

    @property
    def actions(self):
        return 5

    @staticmethod
    def action_name(a):
        return ['noop', 'up', 'down', 'left', 'right'][a]

    @staticmethod
    def from_str(data):
        """Parse an ASCII level (games/pacman.py:28-46) -> (board [1,H,W,4+P] int32 numpy, size, players)."""
        num_players = sum(ch.isdigit() for ch in data)
        lines = data.split('\n')
        board = np.zeros((1, len(lines), len(lines[0]), 4 + num_players), dtype=np.int32)
        for i, line in enumerate(lines):
            for j, ch in enumerate(line):
                layer = {'#': 1, 's': 2, 'S': 3}.get(ch)
                if layer is None:
                    layer = 4 + int(ch) - 1 if ch.isdigit() else 0
                board[0, i, j, layer] = 1
        return board, (len(lines), len(lines[0])), num_players

    # ---- device API -----------------------------------------------------------------------------
    def reset_device(self, data):
        """data [B,H,W,4+P] (any numeric dtype, CUDA or CPU tensor) -> tuple of P observations [B,H,W,4+2P] f32."""
        dev = self.device
        grid = data.detach().to(dev)
        if tuple(grid.shape) != (self.batch_size,) + self.size + (self.grid_depth,):
            raise ValueError('board must have shape %s' % ((self.batch_size,) + self.size + (self.grid_depth,),))
        self._grid = grid.to(torch.int32).contiguous().clone()
        # players in np.where order: lexicographic in (board, x, y, player)   (games/pacman.py:59-61)
        where = (grid[..., self.fields:] == 1).nonzero()
        if where.size(0) != self.batch_size * self.num_players:
            raise ValueError('every board must contain each player exactly once')
        self._players = where.to(torch.int32).contiguous()
        self._scratch = torch.zeros((1,), dtype=torch.int32, device=dev)
        noop = torch.zeros((self.batch_size, self.num_players), dtype=torch.int32, device=dev)
        return self.step_device(noop)[0]

    def step_device(self, actions):
        """actions [B,P] integer tensor -> (tuple of P observations, rewards [P,B] f64 tensor)."""
        dev = self.device
        B, P = self.batch_size, self.num_players
        H, W = self.size
        a = actions.detach().to(device=dev, dtype=torch.int32).contiguous().view(B, P)
        rewards = torch.empty((P, B), dtype=torch.float64, device=dev)
        lib, stream = _lib.lib(), _lib.stream_ptr(dev)
        check(lib.glg_pacman_step(ptr(self._grid), ptr(self._players), ptr(a), ptr(rewards), ptr(self._scratch),
                                  B, H, W, P, stream), 'glg_pacman_step')
        obs = torch.empty((P, B, H, W, self.depth), dtype=torch.float32, device=dev)
        check(lib.glg_pacman_observe(ptr(self._grid), ptr(obs), B, H, W, P, stream), 'glg_pacman_observe')
        return tuple(obs[p] for p in range(P)), rewards

    # ---- reference (numpy) API --------------------------------------------------------------------
    def reset(self, data):
        """`data` [B,H,W,4+P] numpy -> tuple of P numpy observations (games/pacman.py:54-62)."""
        if isinstance(data, torch.Tensor):
            return self.reset_device(data)
        return tuple(o.cpu().numpy() for o in self.reset_device(torch.from_numpy(np.ascontiguousarray(data))))

    def step(self, actions):
        """actions [B,P] numpy -> (tuple of P observations, tuple of P reward vectors [B]) (pacman.py:72-109)."""
        if isinstance(actions, torch.Tensor):
            return self.step_device(actions)
        obs, rewards = self.step_device(torch.from_numpy(np.ascontiguousarray(actions).astype(np.int32)))
        return tuple(o.cpu().numpy() for o in obs), tuple(rewards.cpu().numpy())

    @property
    def grid(self):
        """The reference's [B,H,W,4+2P] int32 numpy grid (last P planes are zeros)."""
        g = self._grid.cpu().numpy()
        return np.concatenate((g, np.zeros(g.shape[:-1] + (self.num_players,), dtype=np.int32)), axis=-1)

    @property
    def players(self):
        return None if self._players is None else self._players.cpu().numpy()

    def copy_board(self):
        return self._grid.cpu().numpy().copy()

    def __repr__(self):
        g = self.grid[0]
        def cell(x, y):
            if g[x, y, 1] == 1:
                return '#'
            if g[x, y, 2] == 1:
                return '•'
            if g[x, y, 3] == 1:
                return '♦'
            for p in range(self.num_players):
                if g[x, y, self.fields + p] == 1:
                    return '{:1d}'.format(p)
            return ' '
        rows = [''.join(cell(x, y) for y in range(self.size[1])) for x in range(self.size[0])]
        bar = '+' + '-' * self.size[0] + '+'
        return bar + '\n' + '\n'.join('|' + r + '|' for r in rows) + '\n' + bar
