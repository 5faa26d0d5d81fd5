"""Golden IMAGES from the REAL reference's drawing code (build container only; needs /root/reference and cv2).

    python tests/golden/make_golden_render.py

Drives games/race.py `tracks_images`, `prettier_tracks` and `record_episode` (its cv2.VideoWriter replaced by a frame
collector) on the first boards of the `iid9` fixture with a fixed action sequence, and writes
tests/golden/render.npz: tracks, actions, the three image sets.  tests/test_render_gpu.py replays the same episode
through game_level_gan_b200 and compares.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

FRAMES = 24          # steps driven; frames 0 (reset's noop), 8, 16 and the last one are kept


def main():
    games = import_reference()
    import cv2
    z = np.load(os.path.join(HERE, 'race_iid9.npz'))
    tracks = torch.from_numpy(z['tracks'][:3])
    cars = [(60., 4., 40.), (80., 2., 60.)]
    with contextlib.redirect_stdout(io.StringIO()):
        env = games.Race(timeout=40., framerate=1. / 20., cars=[games.RaceCar(*c) for c in cars], log_history=True,
                         device=torch.device('cpu'))
    env.record(1)
    env.reset(tracks)
    rng = np.random.default_rng(11)
    acts = rng.choice([1, 1, 1, 4, 7, 0, 2], size=(FRAMES, 2, 3)).astype(np.int64)
    for a in acts:
        env.step(torch.from_numpy(a))
    timgs = env.tracks_images(top_n=3)
    pimgs = env.prettier_tracks(top_n=2, size=320, pad=0.05)

    frames = []

    class Collector(object):
        def __init__(self, *a, **k):
            self.args = a

        def write(self, f):
            frames.append(np.array(f))

        def release(self):
            pass

    real = cv2.VideoWriter
    cv2.VideoWriter = Collector
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            env.record_episode('/tmp/_golden_render_clip')
    finally:
        cv2.VideoWriter = real
    frames = np.stack(frames)
    keep = [0, 8, 16, len(frames) - 1]
    np.savez_compressed(os.path.join(HERE, 'render.npz'), tracks=tracks.numpy(), actions=acts, cars=np.array(cars),
                        record_id=np.int32(1), tracks_images=timgs, prettier=pimgs, frame_ids=np.array(keep),
                        frames=frames[keep], n_frames=np.int32(len(frames)))
    print('tracks_images', timgs.shape, 'prettier', pimgs.shape, 'frames', frames.shape, 'kept', keep)


if __name__ == '__main__':
    main()
