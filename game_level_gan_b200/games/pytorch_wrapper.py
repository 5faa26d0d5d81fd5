"""Adapter from a batched grid environment to PyTorch's [B,C,H,W] layout
(replaces games/pytorch_wrapper.py:6-30).

The reference round-trips every observation through numpy on the host.  For environments that are
`cuda_native` (our Pacman) the wrapper calls their device API instead - same return values (a list of
per-player [B,D,H,W] tensors on the device, permuted views of [B,H,W,D]), no PCIe traffic.  Any other
environment object is driven exactly like the reference does.
"""
import torch


class PytorchWrapper(object):
    def __init__(self, env, device=None):
        self.env = env
        self._wrap_device = device

    def __getattr__(self, attr):
        return getattr(self.env, attr)

    def __repr__(self):
        return self.env.__repr__()

    def _target(self):
        if self._wrap_device is not None:
            return torch.device(self._wrap_device)
        if getattr(self.env, 'cuda_native', False):
            return self.env.device
        return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')

    def _wrap_state(self, state):
        if not isinstance(state, torch.Tensor):
            state = torch.from_numpy(state)
        return state.to(self._target()).permute(0, 3, 1, 2)

    def reset(self, base_board):
        if getattr(self.env, 'cuda_native', False):
            return [self._wrap_state(s) for s in self.env.reset_device(base_board)]
        return [self._wrap_state(s) for s in self.env.reset(base_board.to('cpu').numpy())]

    def step(self, actions):
        if getattr(self.env, 'cuda_native', False) and isinstance(actions, torch.Tensor):
            new_states, rewards = self.env.step_device(actions)
            return [self._wrap_state(s) for s in new_states], rewards
        new_states, rewards = self.env.step(actions)
        return [self._wrap_state(s) for s in new_states], rewards
