"""Benchmark of the batched Race environment step (BASELINE.json metric: race env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = every car of one batch of config 2 of BASELINE.json (4096 synthetic generator-like tracks x 2 cars,
L = 128, 18 ray sensors) advanced by one environment step, its 18-ray observation and reward written = 8192
env-steps.  Actions come from a tape recorded (untimed) with a wall-avoiding heuristic driver, resident in HBM; a
batch is rewound to its mid-race snapshot before every block of <= CYCLE steps and the tape replayed, so ~100 % of
the cars are alive throughout (dead cars skip the ray cast and would inflate the number).

Timed region (`value`, `ms_per_step`, `roofline`): EXACTLY K = --steps steps, issued as ceil(K / CYCLE) calls of
`Race.rollout` (glg_race_rollout, GLG_ROLLOUT_FUSED: one persistent kernel per call, every step's observations and
rewards stored), bracketed by barrier + synchronize on both sides and timed with CUDA events on the launching
stream.  Before the first event the L2 is flushed (a 256 MB device memset, untimed) and consecutive calls rotate
over --replicas independent batches, so the geometry is read from HBM.  The block is repeated `config.repeats`
times (each repeat bracketed and flushed the same way); the MEDIAN block time - max over ranks per repeat - is
reported.  The warm-up runs the identical block (same call shapes, same buffers), at least --warmup steps.

`e2e`    : the same metric through the public API for host-resident callers, host buffers in and out, copies inside
           the timed region: `Race.host_stepper().step` (one CUDA graph per step: H2D, step kernel, D2H, one
           synchronisation per step).
`roofline`: algorithmic bytes (3120 B per track + 145 B per car, per step) x the steps of a launch
           over the launch duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
`parity` : the production path (fused rollout) against the literal kernel on a slice of the bench batch, every run.
`cpu_baseline`: the reference's algorithm on the host cores, ALL 4096 tracks (16 environments of 256, the reference's
           validity check needs 2.7 MB of temporaries per track): the stock reference from baseline/_ref when it is
           there (tools/make_baseline_ref.py), else the torch-op restatement (oracle/race_oracle.py, same ATen
           kernels in the same order); and, for information, the OpenMP C port.
`--impl reference`: times only that CPU arm, same config.
`extra`  : config 4 (2^20 tracks x 4 cars sharded over the ranks, T = 200, winner statistics all-gathered inside
           the timed region), config 3 (episode with two LSTM agents: per-step loop vs CUDA graph) and config 5
           (Pacman step + observation) - informational, N = 1 only for configs 3 and 5.
"""
import argparse
import contextlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_TRACKS, P_CARS, L_SEG, O_RAYS = 4096, 2, 128, 18
REPLICAS = 4     # default of --replicas
CYCLE = 100      # steps between rewinds of a batch (one call of Race.rollout)
PREROLL = 100
SEED = 1234
FLUSH_BYTES = 256 << 20
ALGO_BYTES_PER_TRACK = 3 * (L_SEG + 2) * 8        # centre + left + right points, SURVEY.md 8(d)
ALGO_BYTES_PER_CAR = 145
CONFIG4_CARS = [(60., 4., 40.), (60., 1., 80.), (80., 2., 60.), (50., 3., 50.)]
METRIC = 'race env-steps/sec (envs x players)'
WORKLOAD = ('config2: Race env, 2 players, 4096 synthetic generator-produced tracks, ray-cast sensors')


def synthetic_tracks(n, seed):
    """iid 9-level arcs = the output format of the reference's GeneratorNetworkConvDiscrete."""
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(n, L_SEG, 2)
    t[:, :, 0] = torch.linspace(-1., 1., 9)[torch.randint(0, 9, (n, L_SEG), generator=g)]
    return t


def synthetic_actions(T, P, B, seed, p_forward=0.6):
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(0, 9, (T, P, B), generator=g)
    return torch.where(torch.rand((T, P, B), generator=g) < p_forward, torch.ones_like(a), a)


def driver_actions(states, gen):
    """Wall-avoiding heuristic driver used to record action tapes (not timed): steer towards the side
    with more room, hold a cruising speed of ~0.04-0.06 vmax, 10 % random actions.  Keeps ~100 % of the
    cars alive for hundreds of steps (97 % finish), like the reference's trained agents do."""
    s = states
    left = s[..., 7] + s[..., 8] + 0.5 * s[..., 6]
    right = s[..., 10] + s[..., 11] + 0.5 * s[..., 12]
    steer = torch.zeros(s.shape[:2], dtype=torch.int64, device=s.device)
    steer[right > left + 0.02] = 1
    steer[left > right + 0.02] = 2
    cruise = torch.tensor([0.035, 0.06], device=s.device).repeat((s.size(0) + 1) // 2)[:s.size(0), None]
    thr = torch.where(s[..., 18] > cruise, 0, 1)
    thr = torch.where((s[..., 9] < 0.05) & (s[..., 18] > 0.02), 2, thr)
    a = steer * 3 + thr
    rnd = torch.rand(a.shape, generator=gen, device=s.device) < 0.1
    return torch.where(rnd, torch.randint(0, 9, a.shape, generator=gen, device=s.device), a)


def record_tape(env, tracks, seed, device):
    """reset + PREROLL + CYCLE driver steps; returns (actions [PREROLL+CYCLE,P,B], snapshot at PREROLL)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    states, _ = env.reset(tracks)
    tape, snap = [], None
    for t in range(PREROLL + CYCLE):
        if t == PREROLL:
            snap = env.snapshot()
        a = driver_actions(states, gen)
        tape.append(a)
        states, _ = env.step(a)
    alive_end = float(env.alive.float().mean())
    env.restore(snap)
    return torch.stack(tape), snap, alive_end


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML; `summary()` uses the samples taken between `mark_start()`
    and `mark_end()` (the timed region).  Started before the warm-up so that NVML is initialised by then."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.max_mhz = index, False, [], None
        self.t0 = self.t1 = None
        self.error = None

    def wait_ready(self, timeout=10.):
        """Block until NVML is initialised and the first sample exists (or it failed)."""
        t = time.perf_counter()
        while not self.samples and self.error is None and time.perf_counter() - t < timeout:
            time.sleep(0.001)

    def count_since_start(self):
        return sum(1 for x in self.samples if self.t0 is not None and x[0] >= self.t0)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n for n in dir(nv) if n.startswith('nvmlClocksEventReason') or
                     n.startswith('nvmlClocksThrottleReason')}
            while not self.stop_flag:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                reasons = set()
                for bit, n in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and 'None' not in n and 'All' not in n:
                        reasons.add(n.replace('nvmlClocksEventReason', '').replace('nvmlClocksThrottleReason', ''))
                self.samples.append((time.perf_counter(), mhz, reasons))
                time.sleep(0.001)
        except Exception as e:  # noqa: BLE001
            self.error = 'nvml_unavailable:%s' % type(e).__name__

    def summary(self):
        inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or x[0])]
        if not inside and self.samples and self.t0 is not None:       # region shorter than the sampling period
            inside = [min(self.samples, key=lambda x: abs(x[0] - self.t0))]
        s = sorted(x[1] for x in inside)
        reasons = set()
        for x in inside:
            reasons |= x[2]
        if self.error:
            reasons.add(self.error)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'samples': len(s),
                'reasons': sorted(r for r in reasons if r not in ('GpuIdle', 'ApplicationsClocksSetting'))}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def ncu_traffic():
    path = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
    if os.path.exists(path):
        return json.load(open(path))
    return {}


@contextlib.contextmanager
def stdout_to_stderr():
    """stdout carries ONE JSON line: chatty libraries (NCCL's banner, the reference's import-time prints) write to
    file descriptor 1, which points at stderr inside this block."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def median(xs):
    s = sorted(xs)
    n = len(s)
    return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle/ as the thing timed is allowed only here: cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
CPU_CHUNK = 256      # tracks per CPU environment (the reference's validity check materialises [B, 260^2] temporaries)


def _stock_reference():
    """The UNMODIFIED reference `Race` from baseline/_ref (a copy of /root/reference/{games,utils}, made by
    tools/make_baseline_ref.py, git-ignored), or None.  Its C++ helper needs Boost, which this image lacks: the
    import falls back to the reference's own torch path (IMPL_GPU), which runs on `device=cpu`."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.exists(os.path.join(ref, 'games', 'race.py')):
        return None
    try:
        with stdout_to_stderr():
            sys.path.insert(0, ref)
            import games as ref_games                                  # noqa: F401  (slow: tries to build the helper)
            from games.race import Race as RefRace, RaceCar as RefCar
            sys.path.remove(ref)
        return RefRace, RefCar
    except Exception as e:  # noqa: BLE001
        sys.stderr.write('baseline/_ref could not be imported (%r); timing the restatement instead\n' % (e,))
        return None


def cpu_reference_arm(steps, warmup, n_tracks=None, prefer_stock=True):
    """env-steps/s of the reference's algorithm on the host cores for the bench config: `n_tracks` (default all
    4096) tracks x 2 cars as environments of CPU_CHUNK tracks, each stepped once per step with the same heuristic
    driver the GPU tape was recorded with."""
    n_tracks = B_TRACKS if n_tracks is None else n_tracks
    torch.set_num_threads(os.cpu_count())
    stock = _stock_reference() if prefer_stock else None
    tracks = synthetic_tracks(B_TRACKS, SEED)[:n_tracks]
    gen = torch.Generator().manual_seed(SEED + 1)
    envs, states = [], []
    t_reset = time.perf_counter()
    for lo in range(0, n_tracks, CPU_CHUNK):
        if stock is not None:
            RefRace, RefCar = stock
            env = RefRace(timeout=40., cars=[RefCar(60., 4., 40.), RefCar(60., 1., 80.)], framerate=1. / 20.,
                          log_history=False, device=torch.device('cpu'))
        else:
            from oracle import race_oracle as ro
            env = ro.RaceOracle(timeout=40., cars=ro.default_cars(), framerate=1. / 20.)
        with stdout_to_stderr():
            s, _ = env.reset(tracks[lo:lo + CPU_CHUNK])
        envs.append(env)
        states.append(s)
    t_reset = time.perf_counter() - t_reset

    def step_all():
        for i, env in enumerate(envs):
            states[i], _ = env.step(driver_actions(states[i], gen))

    with torch.no_grad():
        for s in range(warmup):
            step_all()
        t0 = time.perf_counter()
        for s in range(steps):
            step_all()
        dt = time.perf_counter() - t0
    alive = sum(float(e.alive.float().sum()) for e in envs) / (n_tracks * P_CARS)
    kind = 'reference' if stock is not None else 'port'
    what = ('stock reference games/race.py (IMPL_GPU path, device=cpu) from baseline/_ref' if stock is not None else
            'torch-op restatement of games/race.py IMPL_GPU on CPU (oracle/race_oracle.py)')
    return {'value': steps * n_tracks * P_CARS / dt, 'ms_per_step': 1e3 * dt / steps, 'cores': torch.get_num_threads(),
            'kind': kind, 'reset_s': t_reset, 'alive_end': alive, 'same_config': n_tracks == B_TRACKS,
            'sample': '%d of the %d tracks x %d cars (%d environments of %d), %d steps after %d warm-up, %s'
                      % (n_tracks, B_TRACKS, P_CARS, len(envs), CPU_CHUNK, steps, warmup, what)}


def cpu_c_port(steps=10):
    import ctypes
    from game_level_gan_b200.games import _tables
    from oracle import c_oracle
    from oracle import race_oracle as ro
    pr = _tables.race_params(ro.default_cars(), 1. / 20., 40., O_RAYS, 10.)
    cpr = c_oracle.RaceParams()
    ctypes.memmove(ctypes.byref(cpr), ctypes.byref(pr), ctypes.sizeof(cpr))
    env = c_oracle.CRace(cpr)
    st, ct, _ = _tables.heading_tables(L_SEG)
    env.reset(synthetic_tracks(B_TRACKS, SEED).numpy(), st.numpy(), ct.numpy())
    acts = synthetic_actions(steps + 2, P_CARS, B_TRACKS, SEED + 1).numpy()
    env.step(acts[0]); env.step(acts[1])
    t0 = time.perf_counter()
    for s in range(2, steps + 2):
        env.step(acts[s])
    dt = time.perf_counter() - t0
    return {'value': steps * B_TRACKS * P_CARS / dt, 'cores': int(c_oracle.lib().ro_num_threads())}


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_arm(args.steps, args.warmup, n_tracks=args.cpu_tracks)
    line = {'impl': 'reference', 'metric': METRIC, 'value': r['value'],
            'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD + ' (on the host cores)', 'tracks': B_TRACKS, 'players': P_CARS,
                       'segments': L_SEG, 'rays': O_RAYS, 'same_config': r['same_config'],
                       'actions': 'heuristic driver (the one the GPU arm\'s tape is recorded with), computed per step',
                       'alive_fraction_end': r['alive_end'], 'reset_s': r['reset_s']},
            'cpu_baseline': {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': r['kind'],
                             'sample': r['sample']},
            'e2e': {'value': r['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
class Replica(object):
    """One config-2 batch: environment + resident action tape + mid-race snapshot + preallocated outputs."""

    def __init__(self, idx, rank, device, variant, mode):
        from game_level_gan_b200.games import Race, RaceConfig
        self.env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False,
                        device=device, variant=variant)
        seed = SEED + 1000 * rank + idx
        self.tracks = synthetic_tracks(B_TRACKS, seed)
        self.acts, self.snap, self.alive_end = record_tape(self.env, self.tracks, seed + 1, device)
        self.states = torch.empty((CYCLE, P_CARS, B_TRACKS, O_RAYS + 2), dtype=torch.float32, device=device)
        self.rewards = torch.empty((CYCLE, P_CARS, B_TRACKS), dtype=torch.float32, device=device)
        self.mode, self.plans = mode, {}

    def plan(self, n):
        if n not in self.plans:
            self.plans[n] = self.env.rollout_plan(self.acts[PREROLL:PREROLL + n], keep_all=True, mode=self.mode,
                                                  out=(self.states[:n], self.rewards[:n]))
        return self.plans[n]

    def restore(self):
        self.env.restore(self.snap)


def parity_gate(rep, device, tracks_n=256, steps=40):
    """The production path (fused rollout of the packed kernel) against the literal kernel (every ray x every wall,
    one launch per step) on the first `tracks_n` tracks of a bench batch, from reset through PREROLL-like driving:
    every observation, reward and state array compared bit for bit; mismatches are counted, and classified as
    near-boundary when the two readings differ by less than 1e-4 relative (none are expected: the pruning is exact)."""
    from game_level_gan_b200.games import Race, RaceConfig
    tracks = rep.tracks[:tracks_n]
    acts = rep.acts[:steps, :, :tracks_n].contiguous()
    a = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device, variant='fast')
    b = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device, variant='brute')
    sa0, _ = a.reset(tracks)
    sb0, _ = b.reset(tracks)
    sa, ra = a.rollout(acts, keep_all=True, mode='fused')
    sb, rb = b.rollout(acts, keep_all=True, mode='stepwise')

    def diff(x, y):
        x, y = x.float(), y.float()
        bad = ~((x == y) | (torch.isnan(x) & torch.isnan(y)))
        near = bad & ((x - y).abs() <= 1e-4 * torch.maximum(x.abs(), y.abs()).clamp(min=1e-6))
        return int(bad.sum()), int(near.sum())

    pairs = [(sa0, sb0), (sa, sb), (ra, rb), (a.positions, b.positions), (a.directions, b.directions),
             (a.speeds, b.speeds), (a.scores, b.scores), (a.alive, b.alive), (a.finishes, b.finishes),
             (a.winners(), b.winners())]
    bad = near = 0
    for x, y in pairs:
        d = diff(x, y)
        bad += d[0]
        near += d[1]
    return {'car_steps_checked': (steps + 1) * tracks_n * P_CARS, 'values_compared': int(sum(x.numel() for x, _ in pairs)),
            'mismatches': bad, 'near_boundary': near,
            'against': 'literal kernel (GLG_STEP_BRUTE, one launch per step) on the first %d tracks of a bench batch, '
                       'reset + %d driver steps' % (tracks_n, steps)}


def run_config4(args, rank, world, device, barrier, total_tracks, trials=4, T=200):
    """Config 4 of BASELINE.json: 2^20 iid-9 tracks x 4 cars sharded over the ranks by board (all `trials` repetitions
    of a board on one rank, train-gan.py:84), tracks and actions generated ON DEVICE per rank
    (seed + rank; the actions are a tape of the heuristic driver recorded once, untimed), one fused rollout of T = 200 steps, then the per-board winner statistics
    (train-gan.py:103-104) all-gathered - rollout, statistics and collective inside the timed region."""
    import torch.distributed as dist
    from game_level_gan_b200 import dist as gdist
    from game_level_gan_b200.games import Race, RaceCar
    boards = total_tracks // trials
    lo, hi = gdist.shard_bounds(boards, rank, world)
    bl = hi - lo
    B = bl * trials
    P = len(CONFIG4_CARS)
    gen = torch.Generator(device=device).manual_seed(SEED + 4000 + rank)
    levels = torch.randint(0, 9, (bl, L_SEG), generator=gen, device=device, dtype=torch.uint8)
    levels = levels.repeat(trials, 1)                                    # trial-major
    env = Race(timeout=40., cars=[RaceCar(*c) for c in CONFIG4_CARS], framerate=1. / 20., log_history=False, device=device,
               variant=args.variant)
    t0 = time.perf_counter()
    states, _ = env.reset_levels(levels)
    torch.cuda.synchronize()
    reset_s = time.perf_counter() - t0
    # action tape: the heuristic driver plays the episode once, closed loop, on the device (untimed); the timed region
    # replays the tape from the reset state.  (iid random actions would kill every car within ~100 steps, and dead
    # cars skip the ray cast.)
    snap = env.snapshot()
    acts = torch.empty((T, P, B), dtype=torch.int64, device=device)
    for t in range(T):
        acts[t] = driver_actions(states, gen)
        states = env.step(acts[t])[0]
    alive_tape_end = float(env.alive.float().mean())
    del states
    plan = env.rollout_plan(acts, keep_all=False, mode=args.rollout_mode)
    gather = gdist.ShardGather(boards, (P + 1,), torch.float32, device)
    times = []
    for rep in range(3):                                                  # the first repetition is the warm-up
        env.restore(snap)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run()
        stats = gather(env.winner_stats(trials))
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1))
    tm = torch.tensor(times[1:], device=device)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.min())
    fin = gdist.finish_rate(env.finishes)
    alive = float(env.alive.float().mean())
    out = {'workload': 'config4: Race env, 4 players, %d tracks (%d boards x %d trials) sharded over %d GPU(s), T = %d, '
                       'NCCL all-gather of the per-board winner statistics inside the timed region' % (total_tracks, boards, trials, world, T),
           'value': total_tracks * P * T / (ms * 1e-3), 'unit': 'env-steps/s', 'ms': ms, 'tracks_per_gpu': B,
           'reset_s_rank0': reset_s, 'alive_fraction_end_rank0': alive, 'finish_rate': fin,
           'stats_rows': int(stats.size(0)), 'stats_sum': float(stats.sum()),
           'gathered_bytes': int(stats.numel() * 4), 'launches': plan.launches + 2 + (1 if world > 1 else 0),
           'roofline_frac': (ALGO_BYTES_PER_TRACK + ALGO_BYTES_PER_CAR * P) * B * T / (ms * 1e-3) / (measured_peak()[0] * 1e9),
           'alive_fraction_tape_end': alive_tape_end,
           'actions': 'tape of the heuristic driver (closed loop, recorded untimed on the device), replayed from the reset state'}
    del plan, acts, env
    torch.cuda.empty_cache()
    return out


def run_reset(device, variant):
    """`Race.reset` of config 2 (4096 tracks from float [B,L,2] and from 4-bit generator levels): geometry, validity
    (games/race.py:126-211, 326-334), extents, car state and the noop step - device time per call, median of 7."""
    from game_level_gan_b200.games import Race, RaceConfig
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device, variant=variant)
    tracks = synthetic_tracks(B_TRACKS, SEED + 5).to(device)
    levels = torch.round(tracks[:, :, 0] * 4 + 4).to(torch.uint8)
    out = {'workload': 'Race.reset of %d tracks x %d cars (track build + validity + extents + init + noop step)' % (B_TRACKS, P_CARS)}
    for name, call in (('reset_us', lambda: env.reset(tracks)), ('reset_levels_us', lambda: env.reset_levels(levels))):
        times = []
        for i in range(9):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(1e3 * e0.elapsed_time(e1))
        out[name] = median(times)
    out['steps_worth'] = 'one reset = %.0f steps of the fused rollout at 8.6 us' % (out['reset_us'] / 8.6)
    return out


def run_config5(device, T=100):
    """Config 5: Pacman, 65 536 boards of 15x15, 2 players, random actions: step + observation per step."""
    import numpy as np
    from game_level_gan_b200.games import Pacman
    B, H, W, P = 65536, 15, 15, 2
    rng = np.random.default_rng(5)
    fields = rng.choice(4, size=(B, H, W), p=[0.4, 0.5, 0.07, 0.03])          # generators/pacman_generator.py:56
    board = np.zeros((B, H, W, 4 + P), dtype=np.int32)
    np.put_along_axis(board[..., :4], fields[..., None], 1, axis=-1)
    for p, (x, y) in enumerate(((0, 0), (H - 1, W - 1))):                       # :63 opposite corners
        board[:, x, y, :] = 0
        board[:, x, y, 0] = 1
        board[:, x, y, 4 + p] = 1
    env = Pacman((H, W), P, batch_size=B, device=device)
    env.reset_device(torch.from_numpy(board).to(device))
    acts = torch.randint(0, 5, (T, B, P), dtype=torch.int32, device=device)
    for t in range(10):
        env.step_device(acts[t])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(T):
        env.step_device(acts[t])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / T
    # algorithmic bytes: the observation kernel writes P observations of B*H*W*(4+2P) f32 and reads the int32 grid once;
    # the step kernels touch only the players' cells (ncu: 355 MB read + 886 MB written by the observation kernel, which
    # is 211 of the ~236 us - profiles/r02e_pacman_observe_kernel.md)
    obs_bytes = P * B * H * W * (4 + 2 * P) * 4
    grid_bytes = B * H * W * (4 + P) * 4
    peak = measured_peak()[0]
    return {'workload': 'config5: Pacman %d boards of %dx%d, %d players, step + observation kernels' % (B, H, W, P),
            'us_per_step': 1e3 * ms, 'board_steps_per_s': B / (ms * 1e-3),
            'bytes_per_step': obs_bytes + grid_bytes,
            'roofline': {'bound': 'hbm', 'achieved': (obs_bytes + grid_bytes) / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': (obs_bytes + grid_bytes) / (ms * 1e-3) / 1e9 / peak,
                         'note': 'whole step (3 kernels) against observation write P*B*H*W*(4+2P)*4 B + one read of the grid'}}


def run_b200(args, rank, world):
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    from game_level_gan_b200 import dist as gdist
    numa_cores = gdist.bind_host_to_gpu(local) if world > 1 else None      # before any pinned allocation
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group('nccl', device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from game_level_gan_b200 import dist as gdist
    R = max(1, args.replicas)
    reps = [Replica(i, rank, device, args.variant, args.rollout_mode) for i in range(R)]
    alive_start = float(torch.stack([r.env.alive.float().mean() for r in reps]).mean())     # envs sit at their snapshots
    K = args.steps
    chunks = [CYCLE] * (K // CYCLE) + ([K % CYCLE] if K % CYCLE else [])
    flush = torch.empty((FLUSH_BYTES,), dtype=torch.uint8, device=device)
    gather = gdist.ShardGather(world * B_TRACKS, (), torch.int8, device) if world > 1 else None
    launches = [0]

    def block(first):
        """K steps; call j replays the tape of batch (first + j) % R from its snapshot.  The rewind of the first
        batch is done by the caller before the first event."""
        n_launch = 0
        for j, n in enumerate(chunks):
            rep = reps[(first + j) % R]
            if j:
                rep.restore()
            rep.plan(n).run()
            n_launch += rep.plan(n).launches
            if gather is not None and n == CYCLE:     # the one collective of the path, once per episode
                gather(rep.env.winners())
                n_launch += 3
        launches[0] = n_launch

    def timed_block(first):
        barrier()
        reps[first % R].restore()
        flush.zero_()                                  # L2 flush (untimed); also keeps the GPU busy while the host enqueues
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        block(first)
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # warm-up: the identical block (same call shapes, same buffers), at least --warmup steps and every batch once
    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(R, -(-max(args.warmup, 3) // max(K, 1)))
    for i in range(n_warm):
        timed_block(i)
    repeats = args.repeats if args.repeats > 0 else max(3, min(15, 6000 // max(K, 1)))
    sampler.wait_ready()
    sampler.mark_start()
    times = [timed_block(i) for i in range(repeats)]
    # An NVML query takes a few ms and 15 blocks of 20 steps last ~5 ms: keep running the identical blocks (their times
    # join the list the median is taken over, `repeats` grows accordingly) until the sampler has seen the GPU under
    # this load a few times, so that `clocks` describes the timed region and not the idle GPU around it.
    while len(times) < 4000:
        more = torch.tensor([1 if (sampler.error is None and sampler.count_since_start() < 5) else 0], device=device)
        if world > 1:
            dist.all_reduce(more, op=dist.ReduceOp.MAX)         # every rank runs the same number of blocks
        if not int(more.item()):
            break
        times.append(timed_block(len(times)))
    repeats = len(times)
    sampler.mark_end()
    sampler.stop_flag = True
    tm = torch.tensor(times, device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    times = [float(x) for x in tm.tolist()]
    ms = median(times)
    alive_end = sum(r.alive_end for r in reps) / len(reps)     # at the end of a full CYCLE of the tape
    sampler.join(timeout=2)
    env_steps = K * B_TRACKS * P_CARS
    value = world * env_steps / (ms * 1e-3)

    # a lone collective, for the record (N > 1): what one episode end costs
    collective_us = None
    if gather is not None:
        w = reps[0].env.winners()
        for i in range(3):
            gather(w)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            gather(w)
        e1.record()
        barrier()
        collective_us = 1e3 * e0.elapsed_time(e1) / 10

    # ---- end-to-end through the public API with host buffers ----
    from game_level_gan_b200.games import Race, RaceConfig
    e2e_steps = min(max(K, 200), 1000)
    env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device,
               variant=args.variant)
    tape, snap, _ = record_tape(env, synthetic_tracks(B_TRACKS, SEED + 77 + rank), SEED + 78 + rank, device)
    host_acts = tape.cpu().pin_memory()
    stepper = env.host_stepper()          # Race.step for host-resident callers: H2D + kernel + D2H as one CUDA graph per step
    sink = torch.zeros(2)

    def e2e_loop(k):
        for s in range(k):
            if s % CYCLE == 0:
                env.restore(snap)
            st, rw = stepper.step(host_acts[PREROLL + s % CYCLE])     # host actions in, host observations + rewards out
            sink[0] += rw[0, 0]                                        # the host reads the result of every step
        torch.cuda.current_stream().synchronize()

    e2e_loop(10)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(e2e_steps)
    t1.record()
    barrier()
    loop_ms = t0.elapsed_time(t1)
    if world > 1:
        tmx = torch.tensor([loop_ms], device=device)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        loop_ms = float(tmx.item())
    closed_loop_value = world * e2e_steps * B_TRACKS * P_CARS / (loop_ms * 1e-3)

    # the same bytes per step through `Race.host_rollout` (action tape and observations in pinned host memory, chunks
    # of steps pipelined over copy-in / compute / copy-out streams, ONE host synchronisation per call of CYCLE steps)
    hrs = [env.host_rollout(CYCLE, chunk=args.e2e_chunk, mode=args.rollout_mode) for _ in range(2)]
    hr = hrs[0]
    tape_h = host_acts[PREROLL:PREROLL + CYCLE]
    n_calls_e2e = max(10, e2e_steps // CYCLE)

    def e2e_rollouts(k):
        """k calls of CYCLE steps; call i+1 is submitted before the host reads call i's results (two sets of staging
        buffers), every call's results are read by the host before the function returns"""
        pending = None
        for i in range(k):
            env.restore(snap)
            h = hrs[i % 2].submit(tape_h)                              # host tape in (pinned: read by the copy engine in place)
            if pending is not None:
                st, rw = pending.wait()                                # host observations + rewards of every step out
                sink[0] += rw[-1, 0, 0] + st[-1, 0, 0, 0]              # the host reads the results
            pending = h
        st, rw = pending.wait()
        sink[0] += rw[-1, 0, 0] + st[-1, 0, 0, 0]

    e2e_rollouts(2)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_rollouts(n_calls_e2e)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    if world > 1:
        tmx = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmx.item())
    e2e_value = world * n_calls_e2e * CYCLE * B_TRACKS * P_CARS / (e2e_ms * 1e-3)

    extra = {}
    if not args.no_extra:
        extra['config4'] = run_config4(args, rank, world, device, barrier, args.config4_tracks)
    parity = parity_gate(reps[0], device) if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if world == 1 and not args.no_extra:
        extra['config5'] = run_config5(device)
        extra['reset'] = run_reset(device, args.variant)
        extra['config3'] = run_rollout_workload(args, quiet=True)
    peak, peak_kind = measured_peak()
    algo_step = ALGO_BYTES_PER_TRACK * B_TRACKS + ALGO_BYTES_PER_CAR * B_TRACKS * P_CARS
    n_calls = len(chunks)
    step_ms = ms / K
    achieved = algo_step / (step_ms * 1e-3) / 1e9
    fused = reps[0].plan(chunks[0]).launches == 1
    traffic = ncu_traffic().get('race_rollout_fused_kernel' if fused else 'race_step_packed_kernel', {})
    clk = sampler.summary()
    sm_hz = 1e6 * float(clk.get('sm_mhz') or clk.get('sm_max_mhz') or 1965)
    n_smsp = 4 * torch.cuda.get_device_properties(device).multi_processor_count
    ipc = None
    if traffic.get('warp_instructions_per_car_step') and args.variant == 'fast' and B_TRACKS == 4096:
        ipc = traffic['warp_instructions_per_car_step'] * B_TRACKS * P_CARS / (n_smsp * sm_hz * step_ms * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': K, 'warmup': args.warmup, 'ms_per_step': step_ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD + ', per GPU', 'tracks': B_TRACKS, 'players': P_CARS,
                   'segments': L_SEG, 'rays': O_RAYS, 'kernel_variant': args.variant, 'rollout_mode': args.rollout_mode,
                   'repeats': repeats, 'block_ms': times, 'block_ms_median': ms, 'warmup_blocks': n_warm,
                   'l2': 'flushed before every timed block (%d MB device memset, untimed); calls of a block rotate over %d '
                         'independent batches (%.0f MB of geometry each)' % (FLUSH_BYTES >> 20, R, ALGO_BYTES_PER_TRACK * B_TRACKS / 1e6),
                   'alive_fraction': [alive_start, alive_end],
                   'launches': ('%d call(s) of Race.rollout per block of %d steps; ' % (n_calls, K)) +
                               ('one persistent kernel per call (glg_race_rollout, GLG_ROLLOUT_FUSED: track records stay in shared '
                                'memory, car state in registers); ' if fused else 'one step kernel per step; ') +
                               'every step writes its observations and rewards (keep_all)',
                   'state_restore_every_steps': CYCLE, 'actions': 'heuristic-driver tape, race steps %d-%d' % (PREROLL, PREROLL + CYCLE),
                   'collective': None if world == 1 else
                       'all-gather of the winners (int8) at every episode end, i.e. after every full %d-step call '
                       '(%d in a block of %d steps); a lone one takes %.1f us' % (CYCLE, K // CYCLE, K, collective_us),
                   'parallelism': 'dp%d (tracks sharded)' % world,
                   'host_cores_rank0': sorted(numa_cores) if numa_cores else 'unchanged'},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic.get('dram_bytes_per_launch') if traffic.get('steps_per_launch') == (chunks[0] if fused else 1) else None,
                     'traffic_note': traffic.get('source'), 'peak_source': peak_kind,
                     # what actually bounds the kernel: instruction issue.  ncu's instruction count per
                     # car-step x the cars of a step / (SM sub-partitions x clock x measured step time) = issue slots used
                     'issue': None if ipc is None else {
                         'bound': 'instruction issue (1 warp instruction per cycle and SM sub-partition)', 'achieved': ipc,
                         'peak': 1.0, 'unit': 'warp instructions / cycle / sub-partition', 'frac': ipc,
                         'warp_instructions_per_car_step': traffic['warp_instructions_per_car_step'],
                         'sub_partitions': n_smsp, 'sm_mhz': sm_hz / 1e6},
                     'algorithmic_bytes_per_step': algo_step,
                     'algorithmic_bytes_per_launch': algo_step * chunks[0] if fused else algo_step,
                     'steps_per_launch': chunks[0] if fused else 1,
                     'launch_ms': ms / n_calls if fused else step_ms},
        'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'steps': n_calls_e2e * CYCLE,
                'api': 'Race.host_rollout(%d steps, chunk=%d).run(host action tape) -> host observations, rewards of every step '
                       '(pinned memory; chunks pipelined over copy-in / compute / copy-out streams, one host '
                       'synchronisation per call, call i+1 submitted before call i is read; open loop)' % (CYCLE, args.e2e_chunk),
                'h2d_bytes_per_step': hr.h2d_bytes // CYCLE, 'd2h_bytes_per_step': hr.d2h_bytes // CYCLE,
                'ms_per_step': e2e_ms / (n_calls_e2e * CYCLE),
                'closed_loop': {'value': closed_loop_value, 'unit': 'env-steps/s', 'steps': e2e_steps,
                                'api': 'Race.host_stepper().step(host actions) -> host observations, rewards: one CUDA graph '
                                       '(H2D, step kernel, D2H) and one synchronisation PER STEP, what a host-resident '
                                       'policy in the loop pays',
                                'h2d_bytes_per_step': P_CARS * B_TRACKS * 8 + 16,
                                'd2h_bytes_per_step': P_CARS * B_TRACKS * (O_RAYS + 2 + 1) * 4 + 4096}},
        'gpu_launches': launches[0], 'clocks': sampler.summary(), 'parity': parity, 'extra': extra,
    }
    if value < e2e_value:
        line['sanity'] = 'device-resident value below the host round-trip value: the timed region is not steady state'
    if world == 1 and not args.no_cpu:
        tp = cpu_reference_arm(steps=3, warmup=1, n_tracks=args.cpu_tracks)
        cp = cpu_c_port()
        line['cpu_baseline'] = {'value': tp['value'], 'unit': 'env-steps/s', 'cores': tp['cores'], 'kind': tp['kind'],
                                'sample': tp['sample'], 'c_port_value': cp['value'], 'c_port_cores': cp['cores'],
                                'c_port_sample': 'all 4096 tracks x 2 cars, 10 steps, oracle/race_oracle.c (OpenMP)'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# config 3: whole-episode rollout with recurrent policies (informational; also `--workload rollout`)
# ------------------------------------------------------------------------------------------------
class LstmAgents(object):
    """P independent LSTMPolicy(20, 9)-shaped networks (policies/LSTMPolicy.py:6-41: two LSTMCell(256) + linear
    heads, Gumbel-max sampling as in agents/PPOAgent.py:55-63), random weights, recurrent state updated in place."""

    def __init__(self, P, B, device, seed=0):
        torch.manual_seed(seed)
        self.cells = [[torch.nn.LSTMCell(20 if i == 0 else 256, 256).to(device) for i in range(2)] for _ in range(P)]
        self.heads = [torch.nn.Linear(256, 9).to(device) for _ in range(P)]
        self.h = [[(torch.zeros(B, 256, device=device), torch.zeros(B, 256, device=device)) for _ in range(2)]
                  for _ in range(P)]

    def reset(self):
        for per_player in self.h:
            for h, c in per_player:
                h.zero_(); c.zero_()

    def __call__(self, states):
        acts = []
        for p in range(len(self.cells)):
            x = states[p]
            for i, cell in enumerate(self.cells[p]):
                h, c = cell(x, self.h[p][i])
                self.h[p][i][0].copy_(h); self.h[p][i][1].copy_(c)
                x = h
            logits = self.heads[p](x)
            logits[:, 1] += 2.0                                    # untrained nets: bias to "forward" so that cars travel
            gumbel = -torch.log(-torch.log(torch.rand_like(logits).clamp_min(1e-20)))
            acts.append(torch.argmax(logits + gumbel, dim=-1))
        return torch.stack(acts, 0)


def stock_agents(P):
    """The reference's own agents for config 3 - `PPOAgent(9, LSTMPolicy(20, 9))` with the shipped weights
    `learned/agent_{i}_0.pt` (train-gan.py:40-50) - when baseline/_ref holds them (tools/make_baseline_ref.py), else None."""
    root = os.path.join(ROOT, 'baseline', '_ref')
    paths = [os.path.join(root, 'learned', 'agent_%d_0.pt' % i) for i in range(P)]
    if not (os.path.isdir(os.path.join(root, 'agents')) and all(os.path.exists(p) for p in paths)):
        return None
    try:
        with stdout_to_stderr():
            if root not in sys.path:
                sys.path.insert(0, root)
            from agents import PPOAgent
            from policies import LSTMPolicy
            agents = [PPOAgent(9, LSTMPolicy(20, 9)) for _ in range(P)]
            for a, p in zip(agents, paths):
                a.load(p)
        return agents
    except Exception as e:  # noqa: BLE001
        sys.stderr.write('stock agents unavailable (%s: %s), using stand-in LSTM agents\n' % (type(e).__name__, e))
        return None


def run_rollout_workload(args, quiet=False):
    """train-gan.py:86-104 on synthetic boards: reset, play the episode with 2 LSTM agents, winner statistics.
    Compares the reference-style per-step loop with GraphedRollout."""
    from game_level_gan_b200.games import GraphedRollout, Race, RaceConfig
    device = torch.device('cuda', torch.cuda.current_device())
    B, T_limit = 2060, 500                                          # (1024 generated + 6 predefined) x 2 mirrored
    tracks = synthetic_tracks(B, SEED)
    from game_level_gan_b200.games import capture_safe_agents
    out = {'workload': 'config3: episode rollout, %d boards x 2 LSTM(256x2) agents, <= %d steps' % (B, T_limit)}
    with torch.no_grad():
        for mode in ('loop', 'graph'):
            env = Race(timeout=T_limit / 20. - 0.025, cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device)
            stock = stock_agents(2)
            if stock is not None:
                # the reference's loop body, train-gan.py:92 / 95-96 (Gumbel sampling as shipped)
                out['agents'] = 'stock PPOAgent(LSTMPolicy(20, 9)) with the shipped weights learned/agent_{0,1}_0.pt, act(training=False)'
                if mode == 'graph':
                    act, reset_agents = capture_safe_agents(stock)
                else:
                    act = lambda st: torch.stack([a.act(s, training=False) for a, s in zip(stock, st)], dim=0)

                    def reset_agents():
                        for a in stock:
                            a.reset()
            else:
                out['agents'] = 'stand-in LSTMPolicy-shaped networks, random weights (baseline/_ref has no agents)'
                agents = LstmAgents(2, B, device)
                act, reset_agents = agents, agents.reset
            best = None
            for rep in range(2 if quiet else 3):
                states, any_valid = env.reset(tracks)
                reset_agents()
                roll = GraphedRollout(env, act, steps_per_replay=16, on_reset=reset_agents).capture() if mode == 'graph' else None
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if mode == 'loop':
                    while any_valid and not env.finished():
                        states, rewards = env.step(act(states))
                else:
                    roll.run(states)
                stats = env.winner_stats(1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[mode] = {'episode_ms': 1e3 * best, 'steps': env.steps, 'us_per_step': 1e6 * best / max(env.steps - 1, 1),
                         'env_steps_per_s': (env.steps - 1) * B * 2 / best, 'finished_frac': float(env.finishes.float().mean())}
    if not quiet:
        print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--variant', default='fast', choices=['fast', 'warp', 'scan', 'brute'])
    ap.add_argument('--rollout-mode', default='fused', choices=['fused', 'chained', 'stepwise'])
    ap.add_argument('--repeats', type=int, default=0, help='timed blocks (0 = 3..15 depending on --steps); the median is reported')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extra', action='store_true', help='skip the config 3 / 4 / 5 records')
    ap.add_argument('--e2e-chunk', type=int, default=25, help='steps per pipelined chunk of the end-to-end host rollout')
    ap.add_argument('--replicas', type=int, default=REPLICAS, help='independent config-2 batches the calls of a block rotate over')
    ap.add_argument('--tracks', type=int, default=4096, help='tracks per batch (default = config 2)')
    ap.add_argument('--cpu-tracks', type=int, default=None, help='tracks of the CPU arm (default: all, = same config)')
    ap.add_argument('--config4-tracks', type=int, default=1 << 20, help='tracks of the config-4 record over all ranks')
    ap.add_argument('--workload', default='step', choices=['step', 'rollout'],
                    help="'rollout': config 3 alone (episode with LSTM agents), informational, not the contract line")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    globals()['B_TRACKS'] = args.tracks
    if args.workload == 'rollout':
        if rank == 0:
            import __graft_entry__ as entry
            entry.build()
            torch.cuda.set_device(0)
            run_rollout_workload(args)
        return
    if args.impl == 'reference':
        args.steps = min(args.steps, 200)          # bounded: ~0.5 s per step of all 4096 tracks on the host
        args.warmup = min(args.warmup, 5)
        run_reference(args, rank)
        return
    run_b200(args, rank, world)


if __name__ == '__main__':
    main()
