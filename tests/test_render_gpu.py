"""Drawing helpers on a live environment against images made by the REFERENCE's own drawing code for
the same episode (tests/golden/render.npz, made by tests/golden/make_golden_render.py): the history ring, the track
records and the batched ray lengths feed the same rasteriser calls."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip('cv2')


@pytest.fixture(scope='module')
def episode():
    from game_level_gan_b200.games import Race, RaceCar
    z = np.load(os.path.join(GOLDEN, 'render.npz'))
    g = {k: z[k] for k in z.files}
    env = Race(timeout=40., cars=[RaceCar(*c) for c in g['cars'].tolist()], framerate=1. / 20., log_history=True)
    env.record(int(g['record_id']))
    env.reset(torch.from_numpy(g['tracks']))
    for a in g['actions']:
        env.step(torch.from_numpy(a).cuda())
    return env, g


def test_static_images_equal_the_reference(episode):
    env, g = episode
    assert np.array_equal(env.tracks_images(top_n=3), g['tracks_images'])
    assert np.array_equal(env.prettier_tracks(top_n=2, size=320, pad=0.05), g['prettier'])


def test_episode_frames_match_the_reference(episode):
    """Walls, finish line, heading and action arrow come from the same fp32 pixel arithmetic (identical); the green ray
    lengths come from a float evaluation that is not the bit-exact sensor path, so a ray end may land one pixel off:
    at most 0.1 % of the pixels of a frame may differ."""
    from game_level_gan_b200.games import race_render
    env, g = episode
    frames = race_render.episode_frames(env)
    assert frames.shape == (int(g['n_frames']), 480, 640 * 2, 3) and frames.dtype == np.uint8
    for k, f in enumerate(g['frame_ids']):
        diff = (frames[int(f)] != g['frames'][k]).any(-1).mean()
        assert diff <= 1e-3, 'frame %d: %.4f of the pixels differ' % (f, diff)
    # the rays were drawn at all (green pixels around the panel centre)
    green = (frames[0][..., 1] > 120) & (frames[0][..., 0] < 80) & (frames[0][..., 2] < 80)
    assert green.sum() > 200


def test_ray_lengths_agree_with_the_observations(episode):
    """`ray_lengths` (drawing) against the step kernel's own sensor readings of the last step, 1e-4 relative."""
    from game_level_gan_b200.games import race_render
    env, g = episode
    b = int(g['record_id'])
    states, _ = env.step(torch.zeros((2, g['tracks'].shape[0]), dtype=torch.int64).cuda())
    rays = race_render.ray_lengths(env, b, env.positions[b][None].cpu().numpy(), env.directions[b][None].cpu().numpy())[0]
    want = states[:, b, :env.observation_size].cpu() * env.max_distance
    alive = env.alive[b].cpu()
    assert alive.any()
    assert torch.allclose(rays[alive], want[alive], rtol=1e-4, atol=1e-5)


def test_record_episode_writes_a_clip(episode, tmp_path):
    env, g = episode
    out = env.record_episode(str(tmp_path / 'clips' / 'episode'))
    assert out.endswith('.mp4') and os.path.exists(os.path.dirname(out))
    quiet = type(env)(timeout=40., cars=env.cars, framerate=1. / 20., log_history=False)
    assert quiet.record_episode(str(tmp_path / 'none')) is None          # "Logging of history is tuned off."
