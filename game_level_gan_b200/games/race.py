"""Drop-in `Race` environment backed by the sm_100a kernels (replaces games/race.py:9-529, 920-928).

Same constructor, methods, public attributes, shapes and dtypes as the reference class; tensors live
on the environment's CUDA device.  All work of `reset` / `step` / `winners` happens in
game_level_gan_b200/csrc (one fused kernel per step) through the C ABI of include/glg_b200.h.
Semantics are those of the reference's torch path (IMPL_GPU) - the only one that can run without
Boost - including its quirks.  There is no CPU fallback.

Host synchronisation: `step` launches one kernel and returns; the Python control flow the API needs
(`any_valid` from `reset`, `finished()`, the "nobody alive" early-out of `step`) reads a 256-byte
step stamp the kernel maintains, at most once per step and only when asked.
"""
import math

import torch

from .. import _lib
from .._lib import GlgError, RaceState, check, ptr
from . import _tables
from .environment import MultiEnvironment


class RaceCar(object):
    """Car description in the reference's units (games/race.py:9-18)."""

    def __init__(self, max_speed, acceleration, angle):
        # km/h -> units/s (1 unit = 10 m), m/s^2 -> units/s^2, degrees/s -> rad/s
        self.max_speed = max_speed * 100. / 3600.
        self.acceleration = acceleration * 0.1
        self.angle = angle * math.pi / 180.


def _default_device():
    if not torch.cuda.is_available():
        return None
    return torch.device('cuda', torch.cuda.current_device())


class Race(MultiEnvironment):
    IMPL_BOOST = 0
    IMPL_GPU = 1
    IMPL_CPP = 2
    IMPL_B200 = 3

    ACTION_NAMES = ('noop', 'forward', 'backward', 'right', 'forward-right', 'backward-right',
                    'left', 'forward-left', 'backward-left')

    def __init__(self, timeout, cars, observation_size=18, max_distance=10., framerate=1. / 30.,
                 log_history=True, device=None, variant='fast'):
        self._device = torch.device(device) if device is not None else None
        if self._device is not None and self._device.type != 'cuda':
            raise GlgError('Race runs on CUDA devices only (got %s); there is no CPU fallback' % self._device)
        self.timeout = timeout
        self.framerate = framerate
        self.cars = list(cars)
        self.num_players = len(self.cars)
        self.num_tracks = None
        self.steps = 0
        self.steps_limit = int(timeout // framerate)
        self.observation_size = observation_size
        self.max_distance = max_distance
        self.negative_reward = -0.01
        self.log_history = log_history
        self.record_id = 0
        self.game_handle = None
        self._impl_version = Race.IMPL_B200
        self.variant = variant
        self._params = _tables.race_params(self.cars, framerate, timeout, observation_size, max_distance,
                                           step_penalty=self.negative_reward)
        self._const = {}
        self._geom = None
        self._alive_known = None
        self._hist_steps = []
        self._hist = None
        self.positions = self.directions = self.speeds = None
        self._alive = self._finishes = self.scores = None
        self._valid_tracks = None
        self._epoch = 0                 # number of resets (captured CUDA graphs are bound to one episode's buffers)

    # ---- device / constants --------------------------------------------------------------------
    @property
    def device(self):
        if self._device is None:
            self._device = _default_device()
            if self._device is None:
                raise GlgError('no CUDA device available; game_level_gan_b200 has no CPU fallback')
        return self._device

    def _constant(self, name, values):
        if name not in self._const:
            self._const[name] = torch.tensor(values, dtype=torch.float32, device=self.device)
        return self._const[name]

    cars_max_speed = property(lambda self: self._constant('vmax', [c.max_speed for c in self.cars]))
    cars_acceleration = property(lambda self: self._constant('acc', [c.acceleration for c in self.cars]))
    cars_angle = property(lambda self: self._constant('ang', [c.angle for c in self.cars]))
    action_speed = property(lambda self: self._constant('aspeed', [0., 1., -3.] * 3))
    action_dirs = property(lambda self: self._constant('adirs', [0.] * 3 + [1.] * 3 + [-1.] * 3))

    def _heading_tables(self, L):
        key = ('heading', L)
        if key not in self._const:
            s, c, half = _tables.heading_tables(L)
            self._const[key] = (s.to(self.device), c.to(self.device), half)
        return self._const[key]

    def _variant_code(self):
        return {'fast': _lib.STEP_PACKED, 'warp': _lib.STEP_FAST, 'brute': _lib.STEP_BRUTE, 'scan': _lib.STEP_SCAN}[self.variant]

    # ---- reference API -------------------------------------------------------------------------
    def state_shape(self):
        return self.observation_size + 2,

    def players_layer_shape(self):
        pass

    def record(self, board):
        self.record_id = board

    def change_timeout(self, timeout):
        self.timeout = timeout
        self.steps_limit = int(timeout // self.framerate)
        self._params.steps_limit = self.steps_limit

    @property
    def actions(self):
        return 9

    @staticmethod
    def action_name(a):
        return Race.ACTION_NAMES[a]

    @staticmethod
    def pack_levels(levels):
        """[B, L] integer arc levels 0..8 (index into linspace(-1, 1, 9), the discrete generator's output,
        generators/race_track_generator.py:250-261) -> [B, ceil(L/2)] uint8, two levels per byte."""
        lv = levels.to(torch.uint8)
        if lv.size(1) & 1:
            lv = torch.cat((lv, torch.full_like(lv[:, :1], 4)), 1)
        return (lv[:, 0::2] | (lv[:, 1::2] << 4)).contiguous()

    def reset_levels(self, levels):
        """`reset` from the generator's discrete output: `levels` [B, L] integers 0..8 (arc = (level-4)/4,
        width 0).  Same result as reset(tracks) with tracks[..., 0] = linspace(-1, 1, 9)[levels]."""
        levels = levels.to(self.device)
        return self.reset(None, packed_levels=(self.pack_levels(levels), levels.size(1)))

    def reset(self, tracks, geometry=None, packed_levels=None):
        """tracks [B, L, (arc, width)] -> (states [P,B,O+2], any_valid).  games/race.py:116-211.

        `geometry=(centre, left, right)` ([B,L+2,2] each) skips the build and uses the given
        polylines instead (parity tests isolate step parity from build parity this way).
        `packed_levels=(uint8 [B, ceil(L/2)], L)`: see reset_levels.
        """
        dev = self.device
        lib = _lib.lib()
        stream = _lib.stream_ptr(dev)
        with torch.no_grad():
            if packed_levels is not None:
                packed, L = packed_levels
                packed = packed.to(device=dev, dtype=torch.uint8).contiguous()
                B = packed.size(0)
            else:
                tracks = tracks.detach().to(device=dev, dtype=torch.float32).contiguous()
                B, L = tracks.size(0), tracks.size(1)
            N, P = L + 2, self.num_players
            self.steps = 0
            self.num_tracks = B
            self._epoch += 1
            self._lazy = {}
            self._geom = torch.empty((B, 3, N, 2), dtype=torch.float32, device=dev)
            if geometry is not None:
                centre, left, right = (g.to(device=dev, dtype=torch.float32) for g in geometry)
                self._geom[:, 0], self._geom[:, 1], self._geom[:, 2] = right.flip(1), left, centre
            elif packed_levels is not None:
                st, ct, half = self._heading_tables(L)
                check(lib.glg_track_build_levels(ptr(packed), B, L, ptr(st), ptr(ct), half, ptr(self._geom), stream),
                      'glg_track_build_levels')
            else:
                st, ct, half = self._heading_tables(L)
                check(lib.glg_track_build(ptr(tracks), B, L, ptr(st), ptr(ct), half, ptr(self._geom), stream),
                      'glg_track_build')
            self._valid_tracks = torch.empty((B,), dtype=torch.uint8, device=dev)
            check(lib.glg_track_validate(ptr(self._geom), B, N, ptr(self._valid_tracks), stream),
                  'glg_track_validate')
            self._extent = torch.empty((B, 2), dtype=torch.float32, device=dev)
            check(lib.glg_track_extent(ptr(self._geom), B, N, ptr(self._extent), stream), 'glg_track_extent')
            # the six state arrays are views of ONE allocation (snapshot / restore are one copy each)
            K = B * P
            k4 = (K + 3) // 4 * 4
            self._state_buf = torch.empty((6 * k4 * 4 + 2 * k4,), dtype=torch.uint8, device=dev)
            f32 = self._state_buf[:6 * k4 * 4].view(torch.float32)
            self.positions = f32[0:2 * K].view(B, P, 2)
            self.directions = f32[2 * k4:2 * k4 + 2 * K].view(B, P, 2)
            self.speeds = f32[4 * k4:4 * k4 + K].view(B, P)
            self.scores = f32[5 * k4:5 * k4 + K].view(torch.int32).view(B, P)
            self._alive = self._state_buf[6 * k4 * 4:6 * k4 * 4 + K].view(B, P)
            self._finishes = self._state_buf[6 * k4 * 4 + k4:6 * k4 * 4 + k4 + K].view(B, P)
            self._stamp = torch.empty((_lib.ALIVE_SLOTS,), dtype=torch.int32, device=dev)
            self._stamp_host = torch.zeros((_lib.ALIVE_SLOTS,), dtype=torch.int32).pin_memory()
            self._chain = torch.zeros((max(int(lib.glg_race_chain_bytes(B, P)) // 4, 4),), dtype=torch.int32,
                                      device=dev)                                  # per-car stamps / hand-over words (rollout)
            self._seq = 0                                                         # launch sequence number
            self._state = RaceState(ptr(self.positions), ptr(self.directions), ptr(self.speeds),
                                    ptr(self._alive), ptr(self._finishes), ptr(self.scores))
            check(lib.glg_race_init(self._state, B, P, ptr(self._stamp), stream), 'glg_race_init')
            self._stamp_np = self._stamp_host.numpy()      # same memory; numpy compares are cheaper per step
            self._step_const = (lib.glg_race_step, ptr(self._geom), N, ptr(self._valid_tracks), ptr(self._extent),
                                ptr(self._stamp))
            self._alive_known = B * P > 0
            self._hist_steps = []
            self._hist = None
            if self.log_history and B > 0:
                self._hist = torch.zeros((self.steps_limit + 3, P, 6), dtype=torch.float32, device=dev)
            any_valid = bool(self._valid_tracks.any().item()) if B > 0 else False
            noop = torch.zeros((P, B), dtype=torch.int64, device=dev)
            return self.step(noop)[0], any_valid

    def step(self, actions):
        """actions [P,B] int64 -> (states [P,B,O+2] f32, rewards [P,B] f32).  games/race.py:340-500."""
        dev = self.device
        with torch.no_grad():
            B, P, O = self.num_tracks, self.num_players, self.observation_size
            if not (actions.dtype == torch.int64 and actions.device == dev and actions.is_contiguous()):
                actions = actions.detach().to(device=dev, dtype=torch.int64).contiguous()
            if tuple(actions.shape) != (P, B):
                raise ValueError('actions must have shape [num_players, num_boards] = [%d, %d]' % (P, B))
            stream = torch.cuda.current_stream(dev)
            # The reference asks `alive.sum().item() == 0` on every step (a device synchronisation).  Here the answer
            # is only known if the caller asked `finished()` since the last step - which every loop of the reference's
            # scripts does (train-gan.py:91: `while ... not game.finished()`); otherwise the step is enqueued
            # without blocking.  With nobody alive the kernel changes no state either and returns the same rewards; only
            # the width of the returned zeros (19, games/race.py:353-356) is then not reproduced.
            anybody_alive = self._alive_known is not False
            self.steps += 1
            if not anybody_alive:                          # games/race.py:353-356 (19-wide quirk)
                states = torch.zeros((P, B, O + 1), dtype=torch.float32, device=dev)
                rewards = (1. - self.finishes.float()) * self.negative_reward
                return states, rewards.t()
            states = torch.empty((P, B, O + 2), dtype=torch.float32, device=dev)
            rewards = torch.empty((P, B), dtype=torch.float32, device=dev)
            hist = self._history_ring(self.steps)
            if hist is not None:
                self._hist_steps.append(self.steps)
            self._seq += 1
            c = self._step_const                           # pointers that do not change between resets
            rc = c[0](self._params, c[1], B, c[2], actions.data_ptr(), c[3], c[4], self._state, self.steps,
                      states.data_ptr(), rewards.data_ptr(), c[5], self._seq, None, ptr(hist), self.record_id,
                      self._variant_code(), stream.cuda_stream)
            if rc != 0:
                check(rc, 'glg_race_step')
            self._alive_known = None
            return states, rewards

    def step_into(self, actions, states_out, rewards_out, base, offset, stamp=None, history=None):
        """Capture-safe step for CUDA graphs (games/rollout.py): no allocation, no host synchronisation, no
        early-out; the step number and launch number are `offset` plus the two int32 counters in the
        device tensor `base`; steps beyond base[2] do nothing (include/glg_b200.h).  `history`: the device-side
        history ring (`_history_ring(last step)`, grown by the caller BEFORE capture - its address is baked in)."""
        B = self.num_tracks
        check(_lib.lib().glg_race_step(
            self._params, ptr(self._geom), B, self._geom.size(2), ptr(actions), ptr(self._valid_tracks),
            ptr(self._extent), self._state, offset, ptr(states_out), ptr(rewards_out),
            ptr(self._stamp if stamp is None else stamp), offset, ptr(base), ptr(history),
            self.record_id if history is not None else -1, self._variant_code(), _lib.stream_ptr(self.device)), 'glg_race_step')

    def host_stepper(self):
        """A `HostStepper` (games/rollout.py) for this episode: `step` with host-resident actions / observations
        as one CUDA graph per step."""
        from .rollout import HostStepper
        return HostStepper(self)

    def host_rollout(self, T, chunk=25, mode='fused', first_chunk=None):
        """A `HostRollout` (games/rollout.py) for this episode: T-step rollouts whose action tape and observations
        live in pinned host memory, chunked and pipelined over copy-in / compute / copy-out streams
        (`run`, or `submit` / `wait` to overlap consecutive calls)."""
        from .rollout import HostRollout
        return HostRollout(self, T, chunk, mode, first_chunk)

    ROLLOUT_MODES = {'fused': _lib.ROLLOUT_FUSED, 'chained': _lib.ROLLOUT_CHAINED, 'stepwise': _lib.ROLLOUT_STEPWISE}

    def rollout(self, actions, keep_all=False, mode='fused', out=None):
        """T steps with pre-computed actions [T,P,B] (no per-step host work).  Returns the outputs of
        the last step, or of all steps ([T,P,B,O+2], [T,P,B]) with keep_all.  Equivalent to T calls of
        `step` as long as somebody is alive throughout (the 19-wide early-out is not applied).
        `mode`: 'fused' = one persistent kernel for all T steps (track records stay in shared memory, car state in
        registers); 'chained' = one kernel per step, consecutive steps depending on each other car by car;
        'stepwise' = one kernel per step in plain stream order (include/glg_b200.h, glg_race_rollout).  Same results.
        `out=(states, rewards)`: preallocated output tensors of the shapes above."""
        return self.rollout_plan(actions, keep_all=keep_all, mode=mode, out=out).run()

    def rollout_plan(self, actions, keep_all=False, mode='fused', out=None):
        """A `RolloutPlan` for this episode: everything `rollout` prepares (argument checks, output buffers,
        marshalled pointers), done once; `plan.run()` then only enqueues the kernel(s).  For callers that replay
        rollouts of one shape many times (benchmarks, action-tape replay)."""
        return RolloutPlan(self, actions, keep_all, mode, out)

    def _history_ring(self, last_step):
        """The device-side history ring, grown to hold row `last_step` (None when nothing is recorded)."""
        if self._hist is None or not (0 <= self.record_id < self.num_tracks):
            return None
        if last_step >= self._hist.size(0):
            grown = torch.zeros((max(2 * self._hist.size(0), last_step + 1), self.num_players, 6), dtype=torch.float32,
                                device=self.device)
            grown[:self._hist.size(0)] = self._hist
            self._hist = grown
        return self._hist

    def snapshot(self):
        """Copy of the mutable episode state (the reference has no env checkpoint; used to rewind
        rollouts, e.g. by bench.py).  Geometry and validity are not part of it."""
        return {'state': self._state_buf.clone(), 'steps': self.steps,   # _seq is never rewound
                'alive_known': self._any_alive()}

    def restore(self, snap):
        """Rewind to a `snapshot()` of the same episode (device copies only, no host sync)."""
        self._state_buf.copy_(snap['state'], non_blocking=True)
        self.steps = snap['steps']
        self._alive_known = snap['alive_known']

    def _any_alive(self, stream=None):
        if self._alive_known is None:
            self._stamp_host.copy_(self._stamp, non_blocking=True)
            (stream if stream is not None else torch.cuda.current_stream(self.device)).synchronize()
            self._alive_known = bool((self._stamp_np == self._seq).any())
        return self._alive_known

    def finished(self):
        """games/race.py:502-504.  Always refreshes the host's view of "anybody alive" (one 4 KB read + a stream
        synchronisation, what the reference's `alive.sum().item()` costs), which the next `step` uses for the
        reference's early-out."""
        anybody_alive = self._any_alive()
        return self.steps > self.steps_limit or not anybody_alive

    def winners(self):
        """games/race.py:506-529 -> int64 [B] in {-1, 0..P-1}."""
        B, P = self.num_tracks, self.num_players
        out = torch.empty((B,), dtype=torch.int64, device=self.device)
        check(_lib.lib().glg_race_winners(ptr(self.scores), ptr(self._finishes), ptr(self._valid_tracks), B, P,
                                          self.steps_limit, ptr(out), _lib.stream_ptr(self.device)),
              'glg_race_winners')
        return out

    def winner_stats(self, trials):
        """Soft labels for the winner discriminator: one_hot(winners+1, P+1) averaged over `trials`
        repetitions laid out trial-major (train-gan.py:84, 103-104) -> [B/trials, P+1] f32."""
        w = self.winners()
        boards = self.num_tracks // trials
        out = torch.empty((boards, self.num_players + 1), dtype=torch.float32, device=self.device)
        check(_lib.lib().glg_winner_stats(ptr(w), trials, boards, self.num_players, ptr(out),
                                          _lib.stream_ptr(self.device)), 'glg_winner_stats')
        return out

    def iterate_valid(self, agents):
        """games/race.py:336-338."""
        valid = self._valid_tracks.tolist()
        return ((i, a) for i, a in enumerate(agents) if valid[i])

    # ---- reference attribute layouts (views / lazily materialised) ---------------------------------
    alive = property(lambda self: None if self._alive is None else self._alive.view(torch.bool))
    finishes = property(lambda self: None if self._finishes is None else self._finishes.view(torch.bool))

    @property
    def valid(self):
        """bool [B*P], one entry per car (games/race.py:207)."""
        if self._valid_tracks is None:
            return None
        return self._valid_tracks.view(torch.bool).view(-1, 1).repeat(1, self.num_players).view(-1)

    # the track record keeps the right boundary reversed (polyline order, csrc/glg_common.cuh)
    right_vecs = property(lambda self: self._lazy_get('right', lambda: self._geom[:, 0].flip(1)))
    left_vecs = property(lambda self: None if self._geom is None else self._geom[:, 1])
    segments = property(lambda self: None if self._geom is None else self._geom[:, 2])

    def _lazy_get(self, name, fn):
        if self._geom is None:
            return None
        if name not in self._lazy:
            self._lazy[name] = fn()
        return self._lazy[name]

    @property
    def right_bounds(self):     # games/race.py:166
        return self._lazy_get('rb', lambda: torch.cat((self.right_vecs[:, :-1], self.right_vecs[:, 1:]), -1))

    @property
    def left_bounds(self):      # games/race.py:167
        return self._lazy_get('lb', lambda: torch.cat((self.left_vecs[:, :-1], self.left_vecs[:, 1:]), -1))

    def _per_player(self, x):
        return x.unsqueeze(1).repeat(1, self.num_players, 1, 1).view(-1, *x.shape[-2:])

    @property
    def bounds(self):           # games/race.py:168-173, [B*P, 2L+3, 4]
        def make():
            start = torch.cat((self.left_vecs[:, :1], self.right_vecs[:, :1]), -1)
            return self._per_player(torch.cat((self.right_bounds, self.left_bounds, start), 1))
        return self._lazy_get('bounds', make)

    @property
    def reward_bound(self):     # games/race.py:169-171, [B*P, 1, 4]
        return self._lazy_get('rwd', lambda: self._per_player(
            torch.cat((self.left_vecs[:, -1:], self.right_vecs[:, -1:]), -1)))

    @property
    def line_bounds(self):      # games/race.py:175-177, [B*P, 2L+4, 2] on the host
        return self._lazy_get('line', lambda: self._per_player(
            torch.cat((self.right_vecs.flip(1), self.left_vecs), 1)).cpu())

    @property
    def history(self):
        """[(positions, directions, actions, alive)] of board `record_id`, one entry per executed step
        (games/race.py:492-494), materialised from the device-side ring on demand."""
        if self._hist is None or not self._hist_steps:
            return []
        rows = self._hist[torch.tensor(self._hist_steps, device=self._hist.device)].cpu()
        return [(r[:, 0:2].tolist(), r[:, 2:4].tolist(), [int(a) for a in r[:, 4].tolist()],
                 [bool(a) for a in r[:, 5].tolist()]) for r in rows]

    # ---- drawing (games/race.py:531-820).  The rasterisation is OpenCV's, as in the reference; the
    # data path (track records, history ring, batched ray lengths of the recorded cars) is race_render.py ----
    def record_episode(self, filename):
        """games/race.py:531-646: `filename`.mp4 of the recorded board (needs `log_history`)."""
        from . import race_render
        return race_render.record_episode(self, filename)

    def tracks_images(self, top_n=3):
        """games/race.py:648-689 -> uint8 [top_n, 256, 256, 3]."""
        from . import race_render
        return race_render.tracks_images(self, top_n)

    def prettier_tracks(self, top_n=3, size=1024, pad=0.05):
        """games/race.py:691-749 -> uint8 RGBA [top_n, size, size, 4]."""
        from . import race_render
        return race_render.prettier_tracks(self, top_n, size, pad)

    def prettier_tracks_svg(self, top_n=3, size=1024, pad=0.05):
        """games/race.py:751-820 -> [svgwrite.Drawing] (needs the `svgwrite` package, like the reference)."""
        from . import race_render
        return race_render.prettier_tracks_svg(self, top_n, size, pad)


class RolloutPlan(object):
    """Pre-marshalled `glg_race_rollout` call (see `Race.rollout_plan`).  Bound to one episode (one `reset`)."""

    def __init__(self, env, actions, keep_all, mode, out):
        dev = env.device
        if mode not in Race.ROLLOUT_MODES:
            raise ValueError('rollout mode must be one of %s' % sorted(Race.ROLLOUT_MODES))
        self.env, self.epoch = env, env._epoch
        self.actions = actions.detach().to(device=dev, dtype=torch.int64).contiguous()
        B, P, O = env.num_tracks, env.num_players, env.observation_size
        if self.actions.dim() != 3 or tuple(self.actions.shape[1:]) != (P, B):
            raise ValueError('actions must have shape [T, num_players, num_boards] = [T, %d, %d]' % (P, B))
        self.T = T = self.actions.size(0)
        self.keep_all, self.mode = bool(keep_all), mode
        shape_s = (T, P, B, O + 2) if keep_all else (P, B, O + 2)
        shape_r = (T, P, B) if keep_all else (P, B)
        if out is None:
            out = (torch.empty(shape_s, dtype=torch.float32, device=dev), torch.empty(shape_r, dtype=torch.float32, device=dev))
        self.states, self.rewards = out
        for t, shape in ((self.states, shape_s), (self.rewards, shape_r)):
            if tuple(t.shape) != shape or t.dtype != torch.float32 or t.device != dev or not t.is_contiguous():
                raise ValueError('out must be contiguous float32 tensors of shapes %s, %s on %s' % (shape_s, shape_r, dev))
        self._fn = _lib.lib().glg_race_rollout
        self._head = (env._params, ptr(env._geom), B, env._geom.size(2), ptr(self.actions), T, ptr(env._valid_tracks),
                      ptr(env._extent), env._state)
        self._out = (ptr(self.states), ptr(self.rewards), int(self.keep_all), ptr(env._stamp))
        self._chain = ptr(env._chain) if mode == 'chained' else None
        self._tail = (Race.ROLLOUT_MODES[mode], env._variant_code())
        # kernels one call launches (bench.py reports it): the fused mode needs the production kernel's preconditions
        fused = mode == 'fused' and env.variant == 'fast' and O == 18 and env._geom.size(2) <= 256 and env._geom.size(2) % 2 == 0
        self.launches = 1 if fused else T

    def run(self):
        env = self.env
        if env._epoch != self.epoch:
            raise GlgError('the environment was reset after this RolloutPlan was created - make a new one')
        T = self.T
        if T == 0:
            return self.states, self.rewards
        hist = env._history_ring(env.steps + T)
        rc = self._fn(*self._head, env.steps + 1, *self._out, env._seq + 1, self._chain, ptr(hist), env.record_id,
                      *self._tail, torch.cuda.current_stream(env.device).cuda_stream)
        if rc != 0:
            check(rc, 'glg_race_rollout')
        if hist is not None:
            env._hist_steps.extend(range(env.steps + 1, env.steps + T + 1))
        env._seq += T
        env.steps += T
        env._alive_known = None
        return self.states, self.rewards
