"""Drawing helpers of `Race`: `record_episode`, `tracks_images`, `prettier_tracks`,
`prettier_tracks_svg` - the outputs of games/race.py:531-820, produced from what the environment already keeps on the
device: the track records (one copy of the boards drawn, not the per-player `bounds` tensor) and the history ring of
board `record_id`.

Everything is vectorised: world -> pixel transforms are applied to whole coordinate arrays (the reference converts
one Python float at a time), the recorded cars' ray lengths for ALL frames come from one batched ray/wall evaluation
on the device (`ray_lengths`), and OpenCV only rasterises.  Pixel coordinates follow the reference's arithmetic
(truncation towards zero by `int()`, y flipped), so that images agree with the reference's up to anti-aliasing.
"""
import math
import os

import numpy as np
import torch

ACTION_ARROWS = ((0, 0), (0, -15), (0, 15), (15, 0), (10, -10), (-10, 10), (-15, 0), (-10, -10), (10, 10))   # race.py:620-622


def _cv2():
    import cv2
    return cv2


def board_segments(env, board):
    """Walls of one board as float64 [2L+3, 4] = right walls, left walls, start line (the order of `bounds`,
    games/race.py:168-173) and its finish line [4] (`reward_bound`, :169) - one small device -> host copy."""
    geom = env._geom[board].detach().to('cpu', torch.float64).numpy()          # [3, N, 2]: right reversed, left, centre
    right, left = geom[0][::-1], geom[1]
    walls = np.concatenate((np.concatenate((right[:-1], right[1:]), 1), np.concatenate((left[:-1], left[1:]), 1),
                            np.concatenate((left[:1], right[:1]), 1)), 0)
    finish = np.concatenate((left[-1], right[-1]))
    return walls, finish


class Fit(object):
    """The reference's `_move_x / _move_y` (games/race.py:667-671, 711-718): the bounding box of the walls, centred in a
    square of `size` pixels with `pad` (fraction) of margin, y pointing up.  The reference evaluates this with 0-dim
    fp32 tensors and Python scalars, i.e. in fp32 with the constants folded in double first; `pixel` follows that order
    of operations so that the truncation to whole pixels lands on the same side."""

    def __init__(self, walls, size, pad):
        f32 = np.float32
        pts = walls.reshape(-1, 2).astype(f32)
        self.mins, maxs = pts.min(0), pts.max(0)
        self.longer = f32((maxs - self.mins).max())
        self.shift = f32(0.5) * (f32(1.) - (maxs - self.mins) / self.longer)
        self.size, self.pad = size, pad
        self.c0, self.c1 = f32(pad * size), f32((1. - 2. * pad) * size)

    def _u(self, xy):
        xy = np.asarray(xy, dtype=np.float32)
        return self.c0 + self.c1 * ((xy - self.mins) / self.longer + self.shift)

    def real(self, xy):
        """float pixel coordinates (the svg variant, :769-773)"""
        u = self._u(xy).astype(np.float64)
        return np.stack((u[..., 0], self.size - u[..., 1]), -1)

    def pixel(self, xy):
        """integer pixel coordinates: x = int(u), y = size - int(v)"""
        u = np.trunc(self._u(xy)).astype(np.int32)
        return np.stack((u[..., 0], self.size - u[..., 1]), -1)


def _draw_segments(cv2, img, seg_px, colour, thickness):
    """seg_px int32 [n, 4] -> n independent anti-aliased lines (cv2.polylines with 2-point polylines)."""
    cv2.polylines(img, list(seg_px.reshape(-1, 2, 2)), False, colour, thickness=thickness, lineType=cv2.LINE_AA)


def tracks_images(env, top_n=3):
    """uint8 [top_n, 256, 256, 3]: walls in black (2 px), finish line (170, 0, 0) 3 px.  games/race.py:648-689."""
    cv2 = _cv2()
    size = 256
    imgs = 255 * np.ones((top_n, size, size, 3), dtype=np.uint8)
    for i in range(top_n):
        walls, finish = board_segments(env, i)
        fit = Fit(walls, size, 0.05)
        _draw_segments(cv2, imgs[i], fit.pixel(walls.reshape(-1, 2)).reshape(-1, 4), (0, 0, 0, 0), 2)
        _draw_segments(cv2, imgs[i], fit.pixel(finish.reshape(2, 2)).reshape(1, 4), (170, 0, 0, 0), 3)
    return imgs


def _finish_checker(fl, fr, depth):
    """the 2 x 7 chequered finish strip: corner arrays [14, 4, 2] and the dark/light flag per cell (:733-746)"""
    perp = (np.array([fl[1] - fr[1], -(fl[0] - fr[0])]) * depth).astype(np.int32)
    x_steps, y_steps = 2, 7
    xs, ys = perp / x_steps, (fr - fl) / y_steps
    cells, dark = [], []
    for xx in range(x_steps):
        for yy in range(y_steps):
            p = fl + xx * xs + yy * ys
            cells.append(np.stack((p, p + xs, p + ys + xs, p + ys)))
            dark.append((xx % 2 == 0) == (yy % 2 == 0))
    return np.stack(cells), dark


def prettier_tracks(env, top_n=3, size=1024, pad=0.05):
    """uint8 RGBA [top_n, size, size, 4]: tarmac quads in alternating greys (5 segments per band), wall outlines and a
    chequered finish strip on a transparent background.  games/race.py:691-749."""
    cv2 = _cv2()
    imgs = 255 * np.ones((top_n, size, size, 4), dtype=np.uint8)
    imgs[:, :, :, 3] = 0
    for i in range(top_n):
        walls, finish = board_segments(env, i)
        fit = Fit(walls, size, pad)
        n = (walls.shape[0] - 1) // 2
        rpx = fit.pixel(walls[:n].reshape(-1, 2)).reshape(n, 2, 2)             # right walls (p_1, p_2)
        lpx = fit.pixel(walls[n:2 * n].reshape(-1, 2)).reshape(n, 2, 2)        # left walls (l_1, l_2)
        quads = np.stack((lpx[:, 0], lpx[:, 1], rpx[:, 1], rpx[:, 0]), 1)     # l_1, l_2, p_2, p_1
        for j in range(n):
            cv2.fillConvexPoly(imgs[i], quads[j], (70, 70, 70, 255) if (j // 5) % 2 == 0 else (50, 50, 50, 255),
                               lineType=cv2.LINE_AA)
        _draw_segments(cv2, imgs[i], fit.pixel(walls.reshape(-1, 2)).reshape(-1, 4), (5, 5, 5, 255), 1)
        f = fit.pixel(finish.reshape(2, 2))
        cells, dark = _finish_checker(f[0].astype(np.int64), f[1].astype(np.int64), 0.2)
        for c, d in zip(cells, dark):
            cv2.fillConvexPoly(imgs[i], c.astype(np.int32), (20, 20, 20, 255) if d else (210, 210, 210, 255),
                               lineType=cv2.LINE_AA)
    return imgs


def prettier_tracks_svg(env, top_n=3, size=1024, pad=0.05):
    """`svgwrite.Drawing` per board (games/race.py:751-820); needs the `svgwrite` package like the reference."""
    import svgwrite as svg

    def unit(v):
        return v / np.sqrt(np.sum(v ** 2.))

    imgs = [svg.Drawing(shape_rendering='crispEdges') for _ in range(top_n)]
    for i in range(top_n):
        walls, finish = board_segments(env, i)
        fit = Fit(walls, size, pad)
        n = (walls.shape[0] - 1) // 2
        rp = fit.real(walls[:n].reshape(-1, 2)).reshape(n, 2, 2)
        lp = fit.real(walls[n:2 * n].reshape(-1, 2)).reshape(n, 2, 2)
        for j in range(n):
            l1, l2, p2, p1 = lp[j, 0].copy(), lp[j, 1].copy(), rp[j, 1].copy(), rp[j, 0].copy()
            r, u = p2 - l2, l2 - l1
            grow = 0.01 * size                                                 # quads overlap a little: no hairlines
            pts = [l1 + unit(-r - u) * grow, l2 + unit(-r + u) * grow, p2 + unit(r + u) * grow, p1 + unit(r - u) * grow]
            imgs[i].add(imgs[i].polygon(points=[p.tolist() for p in pts],
                                        fill=svg.rgb(70, 70, 70) if (j // 5) % 2 == 0 else svg.rgb(50, 50, 50)))
        f = fit.real(finish.reshape(2, 2))
        fl, fr = f[0].copy(), f[1].copy()
        r = unit(fr - fl) * 0.01 * size
        fl -= r
        fr += r
        cells, dark = _finish_checker(fl, fr, 0.22)
        for c, d in zip(cells, dark):
            imgs[i].add(imgs[i].polygon(points=c.tolist(), fill=svg.rgb(20, 20, 20) if d else svg.rgb(210, 210, 210)))
    return imgs


def ray_lengths(env, board, positions, directions):
    """Sensor readings of recorded cars, clamped to `max_distance`: positions / directions float [F, P, 2] on any device
    -> float32 [F, P, O] on the host.  One batched evaluation on the device of the ray/wall parameter of
    games/race.py:287-308 (t = cross(p - s, w) / cross(d, w), a hit when 0 <= t and the crossing lies on the wall)
    for all frames x players x rays x walls of ONE board; drawing only, not the bit-exact sensor path of the step."""
    dev = env.device
    O = env.observation_size
    geom = env._geom[board]                                                   # [3, N, 2]
    line = torch.cat((geom[0], geom[1]), 0)                                   # polyline right-end ... start ... left-end
    p, q = line[:-1], line[1:]                                                # [W, 2]
    w = q - p
    s = torch.as_tensor(positions, dtype=torch.float32, device=dev)
    nd = torch.as_tensor(directions, dtype=torch.float32, device=dev)
    ang = torch.linspace(-math.pi, math.pi * (1. - 2. / O), O, device=dev)    # race.py:462
    c, sn = torch.cos(ang), torch.sin(ang)
    d = torch.stack((nd[..., 0:1] * c + nd[..., 1:2] * sn, -nd[..., 0:1] * sn + nd[..., 1:2] * c), -1)   # [F, P, O, 2]
    ps = p[None, None, None] - s[:, :, None, None]                            # [F, P, 1, W, 2]
    cross = lambda a, b: a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    den = cross(d[:, :, :, None], w[None, None, None])                        # [F, P, O, W]
    t = cross(ps, w[None, None, None]) / den
    u = cross(ps, d[:, :, :, None]) / den                                     # position of the crossing along the wall
    hit = (t >= 0) & (u >= 0) & (u <= 1) & (den != 0)
    t = torch.where(hit, t, torch.full_like(t, float('inf')))
    return t.min(-1).values.clamp(max=env.max_distance).cpu()


def episode_frames(env):
    """uint8 [frames, 480, 640 * P, 3] (BGR, what cv2.VideoWriter takes): the recorded board seen from every player,
    one 640 x 480 panel per player side by side, cut at the first frame with nobody alive.  games/race.py:531-626."""
    cv2 = _cv2()
    hist = env.history
    if not hist:
        return np.zeros((0, 480, 640 * env.num_players, 3), dtype=np.uint8)
    width, height, scale = 640, 480, 150.
    P = env.num_players
    board = env.record_id
    walls, finish = board_segments(env, board)
    pos = np.array([h[0] for h in hist], dtype=np.float64)                    # [F, P, 2]
    dirs = np.array([h[1] for h in hist], dtype=np.float64)
    acts = np.array([h[2] for h in hist], dtype=np.int64)
    alive = np.array([h[3] for h in hist], dtype=bool)
    dead_frames = np.nonzero(~alive.any(1))[0]
    cut = int(dead_frames[0]) if len(dead_frames) else len(hist)
    rays = ray_lengths(env, board, pos[:cut], dirs[:cut]).numpy().astype(np.float64)             # [cut, P, O]
    O = env.observation_size
    ang = torch.linspace(-math.pi, math.pi * (1. - 2. / O), O).numpy().astype(np.float64)
    record = 255 * np.ones((P, cut, height, width, 3), dtype=np.uint8)
    seg32 = np.concatenate((walls, finish[None]), 0).reshape(-1, 2).astype(np.float32)     # every end point once
    for f in range(cut):
        for pl in range(P):
            img = record[pl, f]
            off = (np.array([width // 2, height // 2], dtype=np.float64) - pos[f, pl] * scale).astype(np.float32)
            px = np.trunc(seg32 * np.float32(scale) + off).astype(np.int32)    # fp32, the reference's order of operations
            px[:, 1] = height - px[:, 1]
            px = px.reshape(-1, 4)
            _draw_segments(cv2, img, px[:-1], (0, 0, 0, 0), 3)
            # rays (green), from the panel centre
            dx, dy = dirs[f, pl]
            rd = np.stack((dx * np.cos(ang) + dy * np.sin(ang), -dx * np.sin(ang) + dy * np.cos(ang)), 1) * (scale * rays[f, pl])[:, None]
            mx, my = width // 2, height // 2
            ends = np.stack((np.trunc(rd[:, 0]).astype(np.int32) + mx, height - np.trunc(rd[:, 1]).astype(np.int32) - my), 1)
            starts = np.tile(np.array([[mx, height - my]], dtype=np.int32), (O, 1))
            _draw_segments(cv2, img, np.concatenate((starts, ends), 1), (0, 170, 0, 0), 1)
            _draw_segments(cv2, img, px[-1:], (170, 0, 0, 0), 3)              # finish line
            hx, hy = int(mx + dx * scale * 0.1), int(my + dy * scale * 0.1)    # heading
            cv2.line(img, (mx, height - my), (hx, height - hy), (100, 149, 237, 0), thickness=4, lineType=cv2.LINE_AA)
            a = int(acts[f, pl])
            if a > 0:
                ox, oy = ACTION_ARROWS[a]
                cv2.arrowedLine(img, (40, height - 40), (2 * ox + 40, height - 40 + 2 * oy), (0, 0, 0, 0), thickness=3,
                                line_type=cv2.LINE_AA)
    frames = np.concatenate(list(record), axis=-2)                             # players side by side
    return np.ascontiguousarray(frames[..., ::-1])                             # RGB -> BGR


def record_episode(env, filename):
    """Writes `filename`.mp4 (MP4V, 1 / framerate fps) with the recorded board's episode.  games/race.py:531-646."""
    if not env.log_history:
        print('Logging of history is tuned off.')
        return None
    cv2 = _cv2()
    frames = episode_frames(env)
    base, _ = os.path.split(filename)
    if base:
        os.makedirs(base, exist_ok=True)
    clip = cv2.VideoWriter(filename + '.mp4', cv2.VideoWriter_fourcc(*'MP4V'), 1. / env.framerate,
                           (640 * env.num_players, 480))
    print()
    print('Saving clip: "{}"'.format(filename + '.mp4'))
    for f in frames:
        clip.write(f)
    print('[{:5d}/{:5d}] ... done.'.format(len(frames), len(frames)))
    clip.release()
    return filename + '.mp4'
