"""B200-native batched Race / Pacman environments behind the reference's `games` API.

Drop-in for the hot path of Grzego/game-level-gan (games/race.py, games/game_helpers.cpp,
games/pytorch_wrapper.py, games/pacman.py): the host side is Python/PyTorch, the work runs in
hand-written sm_100a CUDA kernels reached through the C ABI in include/glg_b200.h.
There is no CPU fallback: importing `game_level_gan_b200.games` works anywhere, but constructing
an environment without the CUDA library or a CUDA device raises.
"""
__version__ = '0.1.0'
