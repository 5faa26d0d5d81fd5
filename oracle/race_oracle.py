"""CPU oracle for the batched Race environment (TEST INFRASTRUCTURE, not product code).

This module restates, in plain torch-on-CPU tensor arithmetic, the algorithm of the
reference's torch path (``IMPL_GPU`` of ``games/race.py``).  It is only ever imported by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs, and only as the checker or the timed CPU baseline - the product (``game_level_gan_b200``)
never imports it and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` (build container only), runs it on seeded inputs and commits the outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement bit-for-bit
against those files.  The reference has no tests / golden vectors of its own.

Why torch and not numpy: the reference's results depend on the exact rounding of ATen's CPU
kernels (``cumsum`` accumulates in double, ``norm`` is sqrt(fma(y,y,x*x)), ``sin``/``cos`` are
SLEEF) - using the same primitive ops is the only way to be bit-identical to it.

Collisions and sensors are evaluated for the cars the reference selects (``update_mask`` / ``alive_mask``);
every arithmetic expression that influences a rounding keeps the reference's operation order.
All line citations are relative to /root/reference/.
"""
import math

import torch

CPU = torch.device('cpu')

# games/race.py:52-71 - per-action throttle and steering flags
THROTTLE = (0., 1., -3., 0., 1., -3., 0., 1., -3.)
STEER = (0., 0., 0., 1., 1., 1., -1., -1., -1.)

SEG_LEN = 0.2        # games/race.py:126
W_MIN, W_MAX = 0.5, 2.0   # games/race.py:127
TURN_DEG = 8.        # games/race.py:140
DRAG = 0.05          # games/race.py:346
STEP_PENALTY = -0.01  # games/race.py:74


class Car(object):
    """games/race.py:9-18 - unit conversion of one car type."""

    def __init__(self, max_speed, acceleration, angle):
        self.max_speed = max_speed * 100. / 3600.
        self.acceleration = acceleration * 0.1
        self.angle = angle * math.pi / 180.


def default_cars():
    """games/race_utils.py:13-16."""
    return [Car(60., 4., 40.), Car(60., 1., 80.)]


# ----------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------

def build_geometry(tracks):
    """Generator output [B, L, (arc, width)] -> centre / left / right polylines [B, L+2, 2].

    Follows games/race.py:135-158 (sentinels, heading cumsum, boundary offsets, centre cumsum).
    """
    tracks = tracks.to(CPU, torch.float32)
    b = tracks.size(0)
    pad = torch.zeros((b, 1, 2))
    t = torch.cat((pad, tracks, pad), dim=1)                         # :136-138
    heading = math.radians(TURN_DEG) * torch.cumsum(t[:, :, :1], dim=1)   # :140
    step = torch.cat((torch.sin(heading), torch.cos(heading)), dim=2) * SEG_LEN  # :142
    normal = torch.stack((step[:, :, 1], -step[:, :, 0]), dim=2)     # :144-145
    off = normal[:, 1:, :] + normal[:, :-1, :]                       # :147
    off = off / off.norm(p=2., dim=-1, keepdim=True)                 # :148
    off = off * (W_MIN + (W_MAX - W_MIN) * t[:, :-1, 1:])            # :149
    first = torch.zeros((b, 1, 2))
    first[:, 0, 0] = W_MIN                                           # :150-151
    off = torch.cat((first, off), dim=1)
    run = torch.cumsum(step, dim=1)                                  # :154
    centre = torch.cat((torch.zeros((b, 1, 2)), run[:, :-1, :]), dim=1)   # :155-156
    right = centre + off                                             # :157
    left = centre + (-off)                                           # :152,158
    return centre, left, right


def wall_table(left, right):
    """Walls in the reference's order [right 0..L, left 0..L, start] plus the finish line.

    games/race.py:166-172.  Returns (walls [B, 2(L+1)+1, 4], finish [B, 1, 4]).
    """
    rw = torch.cat((right[:, :-1, :], right[:, 1:, :]), dim=-1)
    lw = torch.cat((left[:, :-1, :], left[:, 1:, :]), dim=-1)
    start = torch.cat((left[:, :1, :], right[:, :1, :]), dim=-1)
    finish = torch.cat((left[:, -1:, :], right[:, -1:, :]), dim=-1)
    return torch.cat((rw, lw, start), dim=1), finish


def _turn(a, b, c):
    """sign of (b-a).y*(c-b).x - (b-a).x*(c-b).y with a,b [N,S,1,2] and c [N,1,T,2].

    games/race.py:230-238 (three separately rounded fp32 ops: mul, mul, sub).
    """
    ab = b - a
    bc = c - b
    return torch.sign(ab[..., 1] * bc[..., 0] - ab[..., 0] * bc[..., 1])


def _in_box(a, b, c):
    """c inside the axis-aligned box of a,b (games/race.py:240-245)."""
    return ((c <= torch.max(a, b)) & (c >= torch.min(a, b))).all(-1)


def crossing_tables(walls, probes):
    """All wall x probe intersection flags.

    walls [N,S,4], probes [N,T,4] -> (hit [N,S,T], start_on_wall [N,S,T]) exactly as
    games/race.py:248-269 with ``special=True``: ``hit`` excludes the "probe start lies on
    the wall" case, which is returned separately.
    """
    p1, q1 = walls[:, :, None, :2], walls[:, :, None, 2:]
    p2, q2 = probes[:, None, :, :2], probes[:, None, :, 2:]
    o1 = _turn(p1, q1, p2)
    o2 = _turn(p1, q1, q2)
    # the reference evaluates o3/o4 as [N,T,S] and permutes; values are identical
    o3 = _turn(p2.transpose(1, 2), q2.transpose(1, 2), p1.transpose(1, 2)).transpose(1, 2)
    o4 = _turn(p2.transpose(1, 2), q2.transpose(1, 2), q1.transpose(1, 2)).transpose(1, 2)
    hit = (o1 != o2) & (o3 != o4)
    start_on = (o1 == 0.) & _in_box(p1, q1, p2)
    hit = hit | ((o2 == 0.) & _in_box(p1, q1, q2))
    hit = hit | ((o3 == 0.) & _in_box(p2, q2, p1))
    hit = hit | ((o4 == 0.) & _in_box(p2, q2, q1))
    return hit, start_on


def segments_cross(walls, probes):
    """games/race.py:213-269 with ``special=False`` -> bool [N,S,T]."""
    hit, start_on = crossing_tables(walls, probes)
    return hit | start_on


def ray_distances(walls, rays):
    """Smallest ray parameter t per ray, games/race.py:271-308.

    walls [N,S,4]; rays [N,D,(sx,sy,dx,dy)] -> [N,D] (inf where nothing is hit).
    """
    far = rays.clone()
    far[:, :, 2:] = rays[:, :, :2] + 1000. * rays[:, :, 2:]           # :289
    hit, start_on = crossing_tables(walls, far)
    hit = hit & ~start_on                                            # :292
    p, q = walls[:, :, None, :2], walls[:, :, None, 2:]
    s, d = rays[:, None, :, :2], rays[:, None, :, 2:]
    wq = q - p
    ps = p - s
    num = ps[..., 1] * wq[..., 0] - ps[..., 0] * wq[..., 1]           # :300
    den = d[..., 1] * wq[..., 0] - d[..., 0] * wq[..., 1]             # :301
    t = torch.where(hit, num / den, num)
    t = torch.where(start_on, torch.zeros_like(t), t)                # :303
    t = torch.where(hit | start_on, t, torch.full_like(t, float('inf')))   # :305
    t = torch.where(t < 0., torch.full_like(t, float('inf')), t)     # :306
    return torch.min(t, dim=1)[0]                                    # :308 (NaN-propagating)


def tracks_valid(walls, finish, chunk=64):
    """games/race.py:326-334 on cat(walls, finish): no *proper* crossing between any two lines."""
    lines = torch.cat((walls, finish), dim=1)
    out = []
    for i in range(0, lines.size(0), chunk):
        ln = lines[i:i + chunk]
        p1, q1 = ln[:, :, None, :2], ln[:, :, None, 2:]
        p2, q2 = ln[:, None, :, :2], ln[:, None, :, 2:]
        o1 = _turn(p1, q1, p2)
        o2 = _turn(p1, q1, q2)
        o3 = _turn(p2.transpose(1, 2), q2.transpose(1, 2), p1.transpose(1, 2)).transpose(1, 2)
        o4 = _turn(p2.transpose(1, 2), q2.transpose(1, 2), q1.transpose(1, 2)).transpose(1, 2)
        bad = (o1 * o2 < 0) & (o3 * o4 < 0)
        out.append(~bad.flatten(1).any(dim=1))
    return torch.cat(out) if out else torch.zeros((0,), dtype=torch.bool)


def rotate(vecs, angles):
    """games/race.py:310-324: v @ [[cos, -sin], [sin, cos]] (bmm on CPU = mul, mul, add)."""
    m = angles.view(-1, 1).repeat(1, 4).view(-1, 2, 2)
    m[:, 0, 0].cos_()
    m[:, 0, 1].sin_().neg_()
    m[:, 1, 0].sin_()
    m[:, 1, 1].cos_()
    return torch.matmul(vecs.unsqueeze(1), m).view(-1, 2)


def sensor_angles(n):
    """games/race.py:462."""
    return torch.linspace(-math.pi, math.pi * (1. - 2. / n), n)


# ----------------------------------------------------------------------------------------
# environment
# ----------------------------------------------------------------------------------------

class RaceOracle(object):
    """State machine equivalent to reference ``Race`` (IMPL_GPU path) on CPU.

    reset: games/race.py:116-211, step: :340-500, finished: :502-504, winners: :506-529.
    """

    def __init__(self, timeout=40., cars=None, observation_size=18, max_distance=10.,
                 framerate=1. / 20.):
        cars = default_cars() if cars is None else cars
        self.P = len(cars)
        self.vmax = torch.tensor([c.max_speed for c in cars], dtype=torch.float32)
        self.accel = torch.tensor([c.acceleration for c in cars], dtype=torch.float32)
        self.turn = torch.tensor([c.angle for c in cars], dtype=torch.float32)
        self.throttle = torch.tensor(THROTTLE)
        self.steer = torch.tensor(STEER)
        self.O = observation_size
        self.max_distance = max_distance
        self.framerate = framerate
        self.timeout = timeout
        self.steps_limit = int(timeout // framerate)      # games/race.py:47
        self.steps = 0

    # -- reset ---------------------------------------------------------------------------
    def reset(self, tracks, geometry=None):
        """``geometry=(centre,left,right)`` overrides the build (used to isolate step parity)."""
        B = tracks.size(0)
        self.B = B
        self.steps = 0
        if geometry is None:
            geometry = build_geometry(tracks)
        self.centre, self.left, self.right = (g.to(CPU, torch.float32) for g in geometry)
        self.walls, self.finish = wall_table(self.left, self.right)
        P = self.P
        self.pos = torch.zeros((B, P, 2))
        self.pos[:, :, 1] = 0.1                            # :182-183
        self.dir = torch.zeros((B, P, 2))
        self.dir[:, :, 1] = 1.                             # :184-185
        self.speed = torch.zeros((B, P))
        self.alive = torch.ones((B, P), dtype=torch.bool)
        self.finishes = torch.zeros((B, P), dtype=torch.bool)
        self.scores = torch.zeros((B, P), dtype=torch.int32)
        self.valid_tracks = tracks_valid(self.walls, self.finish)      # :200
        self.valid = self.valid_tracks.view(-1, 1).repeat(1, P).view(-1)   # :207
        any_valid = bool(self.valid_tracks.any())
        states, _ = self.step(torch.zeros((P, B), dtype=torch.int64))  # :210
        return states, any_valid

    # -- step ----------------------------------------------------------------------------
    def step(self, actions):
        B, P, O = self.B, self.P, self.O
        act = actions.to(CPU).t().contiguous().clone()
        self.steps += 1
        if int(self.alive.sum()) == 0:                     # :353-356 (19-wide quirk)
            states = torch.zeros((P, B, O + 1))
            rewards = (1. - self.finishes.float()) * STEP_PENALTY
            return states, rewards.t()
        valid = self.valid.view(B, P)
        act[~self.alive | ~valid] = 0                      # :359
        flat = act.view(-1)
        # heading (:362-364)
        ang = self.framerate * self.steer[flat] * self.turn.repeat(B).view(-1)
        ndir = rotate(self.dir.view(-1, 2), ang).view(B, P, 2)
        # speed (:366-370)
        thr = self.throttle[flat].view(B, P)
        v = self.speed + self.framerate * thr * self.accel[None, :]
        nv = torch.min(self.vmax, v.clamp(min=0.))
        moving = (torch.abs(nv) > 1e-7)
        npos = self.pos + ndir * nv[:, :, None]            # :372
        # progress (:374-376; divisor is bounds.size(2)-1 == 3)
        gap = torch.norm(npos[:, :, None, :] - self.centre[:, None, :, :], dim=-1)
        idx = gap.argmin(-1)
        prog = idx.float() / 3
        upd = self.alive & moving & valid                  # :380
        rewards = torch.zeros((B, P))
        rewards[~self.finishes] = STEP_PENALTY             # :382-383
        # collisions (:385-432): like the reference, only for the cars selected by `upd` (update_mask, :380)
        paths = torch.cat((self.pos, npos), dim=-1)        # [B,P,4]
        dead = torch.zeros((B * P,), dtype=torch.bool)
        done = torch.zeros((B * P,), dtype=torch.bool)
        sel_u = upd.view(-1).nonzero().squeeze(-1)
        if sel_u.numel() > 0:                              # :385
            car_track = sel_u // P
            pth = paths.view(-1, 1, 4)[sel_u]              # :390
            dead[sel_u] = segments_cross(self.walls[car_track], pth).any(dim=1).squeeze(-1)     # :406-407
            done[sel_u] = segments_cross(self.finish[car_track], pth).any(dim=1).squeeze(-1)    # :431-432
        dead, done = dead.view(B, P), done.view(B, P)
        self.last_dead, self.last_done, self.last_idx = dead, done, idx
        rewards = torch.where(upd, rewards + (done.float() - dead.float()), rewards)   # :434
        self.alive = self.alive & ~dead & ~done            # :414,435
        self.finishes = self.finishes | done               # :436
        sc = torch.where(dead, idx.int() + (self.steps_limit + 1), self.scores)        # :442-444
        self.scores = torch.where(done, torch.full_like(sc, self.steps), sc)           # :446-447
        nv = torch.where(self.alive, nv, torch.zeros_like(nv))                         # :449
        drags = 1. - (1. - (thr != 0.).float()) * DRAG     # :452
        self.dir = ndir
        self.speed = nv * drags
        self.pos = npos
        # sensors (:459-489) for cars alive after the update
        obs = torch.zeros((B * P, O))
        am = self.alive.view(-1)
        if bool(am.any()):
            oa = sensor_angles(O).view(1, -1).repeat(B * P, 1).view(-1)
            rd = ndir.reshape(-1, 1, 2).repeat(1, O, 1).view(-1, 2)
            od = rotate(rd, oa).view(-1, O, 2)
            rays = torch.cat((npos.reshape(-1, 1, 2).repeat(1, O, 1), od), dim=-1)
            walls_pc = self.walls.unsqueeze(1).expand(-1, P, -1, -1).reshape(B * P, -1, 4)
            sel = am.nonzero().squeeze(-1)
            dist = _chunked(ray_distances, walls_pc, rays, sel)
            obs[sel] = dist.clamp(max=self.max_distance) / self.max_distance           # :489
        states = torch.cat((obs.view(B, P, O),
                            (self.speed / self.vmax[None, :])[:, :, None],
                            prog[:, :, None]), dim=-1)     # :496-499
        return states.permute(1, 0, 2), rewards.t()

    def finished(self):
        return self.steps > self.steps_limit or not bool(self.alive.any())

    def winners(self):
        """games/race.py:506-529."""
        big = self.steps_limit + 1
        anyf = self.finishes.any(dim=-1)
        fin = torch.where(self.finishes, self.scores, torch.full_like(self.scores, big))
        w = torch.where(anyf, fin.argmin(dim=-1), self.scores.argmax(dim=-1))
        w[~self.valid_tracks] = -1
        return w


def _chunked(fn, walls, rays, sel, chunk=2048):
    out = []
    for i in range(0, sel.numel(), chunk):
        s = sel[i:i + chunk]
        out.append(fn(walls[s], rays[s]))
    return torch.cat(out)
