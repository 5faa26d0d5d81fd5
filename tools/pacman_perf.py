"""Config 5: Pacman, 65 536 boards of 15x15 with 2 players - device-resident step + observation throughput."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from game_level_gan_b200.games import Pacman
dev = torch.device('cuda', 0); torch.cuda.set_device(0)
B, H, W, P, T = 65536, 15, 15, 2, 100
rng = np.random.default_rng(5)
fields = rng.choice(4, size=(B, H, W), p=[0.4, 0.5, 0.07, 0.03])
board = np.zeros((B, H, W, 4 + P), dtype=np.int32)
np.put_along_axis(board[..., :4], fields[..., None], 1, axis=-1)
for p, (x, y) in enumerate(((0, 0), (H - 1, W - 1))):
    board[:, x, y, :] = 0; board[:, x, y, 0] = 1; board[:, x, y, 4 + p] = 1
env = Pacman((H, W), P, batch_size=B)
env.reset_device(torch.from_numpy(board).to(dev))
acts = torch.randint(0, 5, (T, B, P), dtype=torch.int32, device=dev)
for t in range(10): env.step_device(acts[t])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(T): env.step_device(acts[t])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / T
obs_bytes = P * B * H * W * (4 + 2 * P) * 4
grid_bytes = 2 * B * H * W * (4 + P) * 4
print('pacman B=%d %dx%d P=%d: %.1f us per step, %.3e board-steps/s, obs write %.0f MB + grid r/w %.0f MB per step -> %.0f GB/s (%.2f of 6549.8)' % (
    B, H, W, P, 1e3 * ms, B / (ms * 1e-3), obs_bytes / 1e6, grid_bytes / 1e6, (obs_bytes + grid_bytes) / (ms * 1e-3) / 1e9, (obs_bytes + grid_bytes) / (ms * 1e-3) / 6549.8e9))

# the two halves separately
from game_level_gan_b200 import _lib
from game_level_gan_b200._lib import ptr
lib = _lib.lib(); stream = _lib.stream_ptr(dev)
obs = torch.empty((P, B, H, W, 4 + 2 * P), dtype=torch.float32, device=dev)
rewards = torch.empty((P, B), dtype=torch.float64, device=dev)
for name, fn in (('observe kernel', lambda t: lib.glg_pacman_observe(ptr(env._grid), ptr(obs), B, H, W, P, stream)),
                 ('step kernels', lambda t: lib.glg_pacman_step(ptr(env._grid), ptr(env._players), ptr(acts[t]), ptr(rewards), ptr(env._scratch), B, H, W, P, stream))):
    for t in range(5): fn(t)
    torch.cuda.synchronize(); e0.record()
    for t in range(T): fn(t)
    e1.record(); torch.cuda.synchronize()
    print('%s: %.1f us' % (name, 1e3 * e0.elapsed_time(e1) / T))
