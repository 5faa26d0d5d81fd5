"""Benchmark of the batched Race environment step (BASELINE.json metric: race env-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one launch of the fused step kernel over one batch of config 2 of BASELINE.json
(4096 synthetic generator-like tracks x 2 cars, L = 128, 18 ray sensors) = 8192 env-steps.
R independent replicas of that batch are stepped round-robin so that the geometry touched between two
visits of a replica (R x 12.8 MB) exceeds the 126 MB L2 - no L2 flush kernels in the timed region.
Actions come from a tape recorded (untimed) with a wall-avoiding heuristic driver, resident in HBM; every
CYCLE steps a replica's car state is rewound to its mid-race snapshot (6 small device copies, inside the
timed region) and the same tape is replayed, so ~100 % of the cars are alive throughout - dead cars skip
the ray cast and would inflate the number.  The alive fraction at both ends of the cycle is reported.

`value`  : device-resident throughput, CUDA events around the K launches (max over ranks); the launches of a
           cycle are one `Race.rollout` (glg_race_rollout: one kernel per step, consecutive steps chained car by
           car), and every step writes its observations and rewards (keep_all).
`e2e`    : the same metric through the public API for host-resident callers (`Race.host_stepper().step`):
           host actions in, host observations + rewards out EVERY step (H2D, step kernel, D2H as one CUDA graph,
           one synchronisation per step), inside the timed region.
`roofline`: algorithmic bytes of one launch (3120 B per track + 145 B per car) over the
           average launch duration, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline`: the reference's algorithm on the host cores - the torch-op restatement (same ATen
           kernels as the reference's IMPL_GPU path on CPU) and, for information, the OpenMP C port.
`--impl reference`: times only that CPU restatement (the reference itself cannot travel to the GPU box;
           its C++ helper needs Boost, which this image lacks - DESIGN.md section 7).
"""
import argparse
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
REPLICAS = 16   # default of --replicas
CYCLE = 100
KEEP_ALL = True      # every step's observations and rewards are kept ([CYCLE,P,B,20] per cycle), not only the last
PREROLL = 100
SEED = 1234
ALGO_BYTES_PER_TRACK = 3 * (L_SEG + 2) * 8        # centre + left + right points, SURVEY.md 8(d)
ALGO_BYTES_PER_CAR = 145


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
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n for n in dir(nv) if n.startswith('nvmlClocksEventReason') or
                     n.startswith('nvmlClocksThrottleReason')}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, n in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and 'None' not in n and 'All' not in n:
                        self.reasons.add(n.replace('nvmlClocksEventReason', '').replace('nvmlClocksThrottleReason', ''))
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.reasons.add('nvml_unavailable:%s' % type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(r for r in self.reasons if r not in ('GpuIdle', 'ApplicationsClocksSetting'))}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def ncu_traffic():
    path = os.path.join(ROOT, 'profiles', 'step_kernel_traffic.json')
    if os.path.exists(path):
        return json.load(open(path)).get('dram_bytes_per_launch')
    return None


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle/ as the thing timed is allowed only here: cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_torch_port(steps, warmup, n_tracks=256):
    from oracle import race_oracle as ro
    torch.set_num_threads(os.cpu_count())
    env = ro.RaceOracle(timeout=40., cars=ro.default_cars(), framerate=1. / 20.)
    tracks = synthetic_tracks(B_TRACKS, SEED)[:n_tracks]
    gen = torch.Generator().manual_seed(SEED + 1)
    states, _ = env.reset(tracks)
    for s in range(warmup):
        states, _ = env.step(driver_actions(states, gen))
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        states, _ = env.step(driver_actions(states, gen))
    dt = time.perf_counter() - t0
    return {'value': steps * n_tracks * P_CARS / dt, 'ms_per_step': 1e3 * dt / steps, 'cores': torch.get_num_threads(),
            'sample': '%d of the %d tracks x %d cars, %d steps after %d warm-up, torch-op restatement of '
                      'games/race.py IMPL_GPU on CPU' % (n_tracks, B_TRACKS, P_CARS, steps, warmup),
            'alive_end': float(env.alive.float().mean())}


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
    r = cpu_torch_port(args.steps, args.warmup)
    line = {'impl': 'reference', 'metric': 'race env-steps/sec (envs x players)', 'value': r['value'],
            'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'config2: Race env, 2 players, 4096 synthetic generator-produced tracks, '
                                   'ray-cast sensors (bounded sample per step, see cpu_baseline.sample)'},
            'cpu_baseline': {'value': r['value'], 'unit': 'env-steps/s', 'cores': r['cores'], 'kind': 'port',
                             'sample': r['sample']},
            'e2e': {'value': r['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
class Replica(object):
    """One config-2 batch: environment + resident actions + mid-race snapshot."""

    def __init__(self, idx, rank, device, variant):
        from game_level_gan_b200.games import Race, RaceConfig
        self.env = Race(timeout=40., cars=RaceConfig.cars, framerate=1. / 20., log_history=False,
                        device=device, variant=variant)
        seed = SEED + 1000 * rank + idx
        self.acts, self.snap, self.alive_end = record_tape(self.env, synthetic_tracks(B_TRACKS, seed), seed + 1, device)

    def restore(self):
        self.env.restore(self.snap)

    def cycle(self, n):
        self.restore()
        self.env.rollout(self.acts[PREROLL:PREROLL + n], keep_all=KEEP_ALL)


def run_b200(args, rank, world):
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        # stdout carries ONE JSON line: NCCL prints its version banner to stdout when the communicator is created,
        # so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    REPLICAS = args.replicas
    reps = [Replica(i, rank, device, args.variant) for i in range(REPLICAS)]
    alive_start = float(torch.stack([r.env.alive.float().mean() for r in reps]).mean())     # envs sit at their snapshots

    def run_steps(k):
        done, i = 0, 0
        while done < k:
            n = min(CYCLE, k - done)
            reps[i % REPLICAS].cycle(n)
            done += n
            i += 1

    run_steps(max(args.warmup, 3))
    if world > 1:       # warm-up of the collective too (communicator set-up is not part of a step)
        from game_level_gan_b200 import dist as gdist
        gdist.all_gather_winners(reps[0].env.winners())
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps)
    if world > 1:       # the one collective of the path: winners of every shard
        from game_level_gan_b200 import dist as gdist
        gdist.all_gather_winners(reps[0].env.winners())
    e1.record()
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    alive_end = sum(r.alive_end for r in reps) / len(reps)     # at the end of a full CYCLE of the tape
    if world > 1:
        tm = torch.tensor([ms], device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    sampler.join(timeout=2)
    env_steps = args.steps * B_TRACKS * P_CARS
    value = world * env_steps / (ms * 1e-3)

    # ---- end-to-end through the public API with host buffers ----
    from game_level_gan_b200.games import Race, RaceConfig
    e2e_steps = min(args.steps, 400)
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
    e2e_ms = t0.elapsed_time(t1)
    if world > 1:
        tm = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_ms = float(tm.item())
    e2e_value = world * e2e_steps * B_TRACKS * P_CARS / (e2e_ms * 1e-3)

    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    algo_bytes = ALGO_BYTES_PER_TRACK * B_TRACKS + ALGO_BYTES_PER_CAR * B_TRACKS * P_CARS
    launch_ms = ms / args.steps
    achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
    line = {
        'metric': 'race env-steps/sec (envs x players)', 'value': value, 'unit': 'env-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': launch_ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config2: Race env, 2 players, 4096 synthetic generator-produced tracks, '
                               'ray-cast sensors, per GPU', 'tracks': B_TRACKS, 'players': P_CARS,
                   'segments': L_SEG, 'rays': O_RAYS, 'kernel_variant': args.variant,
                   'l2': '%d replicas stepped round-robin, %.0f MB of geometry > 126 MB L2 (no flush)'
                         % (REPLICAS, REPLICAS * ALGO_BYTES_PER_TRACK * B_TRACKS / 1e6),
                   'alive_fraction': [alive_start, alive_end],
                   'launches': 'one step kernel per step (glg_race_rollout: chained car by car, every step writes its '
                               'observations and rewards, keep_all=%s)' % KEEP_ALL,
                   'state_restore_every_steps': CYCLE, 'actions': 'heuristic-driver tape, race steps %d-%d' % (PREROLL, PREROLL + CYCLE), 'parallelism': 'dp%d (tracks sharded)' % world},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': ncu_traffic(), 'peak_source': peak_kind,
                     'algorithmic_bytes_per_launch': algo_bytes},
        'e2e': {'value': e2e_value, 'unit': 'env-steps/s', 'steps': e2e_steps,
                'api': 'Race.host_stepper().step(host actions) -> host observations, rewards (one CUDA graph per step)',
                'h2d_bytes_per_step': P_CARS * B_TRACKS * 8 + 12,
                'd2h_bytes_per_step': P_CARS * B_TRACKS * (O_RAYS + 2 + 1) * 4 + 4096},
        'gpu_launches': args.steps, 'clocks': sampler.summary(),
    }
    if world == 1 and not args.no_cpu:
        tp = cpu_torch_port(steps=20, warmup=3)
        cp = cpu_c_port()
        line['cpu_baseline'] = {'value': tp['value'], 'unit': 'env-steps/s', 'cores': tp['cores'], 'kind': 'port',
                                'sample': tp['sample'], 'c_port_value': cp['value'], 'c_port_cores': cp['cores'],
                                'c_port_sample': 'all 4096 tracks x 2 cars, 10 steps, oracle/race_oracle.c (OpenMP)'}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# config 3: whole-episode rollout with recurrent policies (informational, `--workload rollout`)
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


def run_rollout_workload(args):
    """train-gan.py:86-104 on synthetic boards: reset, play the episode with 2 LSTM agents, winner statistics.
    Prints one JSON line comparing the reference-style per-step loop with GraphedRollout."""
    from game_level_gan_b200.games import GraphedRollout, Race, RaceConfig
    device = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    B, T_limit = 2060, 500                                          # (1024 generated + 6 predefined) x 2 mirrored
    tracks = synthetic_tracks(B, SEED)
    out = {'workload': 'config3: episode rollout, %d boards x 2 LSTM(256x2) agents, <= %d steps' % (B, T_limit)}
    with torch.no_grad():
        for mode in ('loop', 'graph'):
            env = Race(timeout=T_limit / 20. - 0.025, cars=RaceConfig.cars, framerate=1. / 20., log_history=False, device=device)
            agents = LstmAgents(2, B, device)
            best = None
            for rep in range(3):
                states, any_valid = env.reset(tracks)
                agents.reset()
                roll = GraphedRollout(env, agents, steps_per_replay=16, on_reset=agents.reset).capture() if mode == 'graph' else None
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                if mode == 'loop':
                    while any_valid and not env.finished():
                        states, rewards = env.step(agents(states))
                else:
                    roll.run(states)
                stats = env.winner_stats(1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            out[mode] = {'episode_ms': 1e3 * best, 'steps': env.steps, 'us_per_step': 1e6 * best / max(env.steps - 1, 1),
                         'env_steps_per_s': (env.steps - 1) * B * 2 / best, 'finished_frac': float(env.finishes.float().mean())}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--variant', default='fast', choices=['fast', 'warp', 'scan', 'brute'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--replicas', type=int, default=REPLICAS, help='independent config-2 batches stepped round-robin')
    ap.add_argument('--tracks', type=int, default=4096, help='tracks per batch (default = config 2)')
    ap.add_argument('--workload', default='step', choices=['step', 'rollout'],
                    help="'rollout': config 3 (episode with LSTM agents), informational, not the contract line")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    globals()['B_TRACKS'] = args.tracks
    if args.workload == 'rollout':
        if rank == 0:
            import __graft_entry__ as entry
            entry.build()
            run_rollout_workload(args)
        return
    if args.impl == 'reference':
        if args.steps > 400:
            args.steps = 200          # bounded: ~0.1 s per sampled step on the host
        args.warmup = min(args.warmup, 5)
        run_reference(args, rank)
        return
    run_b200(args, rank, world)


if __name__ == '__main__':
    main()
