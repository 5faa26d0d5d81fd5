"""Whole-episode rollouts without per-step host work.

The reference plays an episode with a Python loop that synchronises with the device twice per step
(train-gan.py:91-93: `while any_valid and not game.finished(): actions = ...act(s)...; game.step(actions)`).
`GraphedRollout` captures `k` iterations of [policy -> env step] in ONE CUDA graph - the policy stays
whatever PyTorch code the caller provides (the reference's PPOAgent / LSTMPolicy), the step is the
sm_100a kernel - replays it, and looks at `finished()` once per replay.  Steps that a replay runs after
the episode has ended change nothing (dead and finished cars are frozen, steps past the time limit are
no-ops in the kernel), and `Race.steps` is set to what the reference's loop would have counted, so the
state of the environment after `run()` is the same as after the reference's loop.
"""
import numpy as np
import torch

from .. import _lib
from .._lib import GlgError


class GraphedRollout(object):
    def __init__(self, env, act, steps_per_replay=8, on_reset=None):
        """env: a reset `Race`; act(states [P,B,O+2]) -> actions [P,B] int64 (capture-safe torch code:
        no host synchronisation, no data-dependent shapes; recurrent state kept in tensors updated in
        place).  on_reset(): called after the warm-up pass of `capture` and at the start of `run` (e.g.
        zero the policies' recurrent state, like PPOAgent.reset, agents/PPOAgent.py:40-44)."""
        if env.num_tracks is None or env.num_tracks == 0:
            raise GlgError('GraphedRollout needs a reset environment with at least one track')
        self.env, self.act, self.k, self.on_reset = env, act, int(steps_per_replay), on_reset
        self.epoch = env._epoch
        dev = env.device
        B, P, O = env.num_tracks, env.num_players, env.observation_size
        self.states = torch.zeros((P, B, O + 2), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((P, B), dtype=torch.float32, device=dev)
        self.actions = torch.zeros((P, B), dtype=torch.int64, device=dev)
        self.base = torch.zeros((3,), dtype=torch.int32, device=dev)      # {step number, launch number} so far, last step
        # history of the recorded board (games/race.py:492-494): rows up to the last step a replay can execute
        self.hist = env._history_ring(env.steps_limit + 1 + self.k)
        self.graph = None

    def _body(self):
        for j in range(1, self.k + 1):
            self.actions.copy_(self.act(self.states))
            self.env.step_into(self.actions, self.states, self.rewards, self.base, j, history=self.hist)
        self.base[:2] += self.k

    def capture(self):
        side = torch.cuda.Stream(device=self.env.device)
        side.wait_stream(torch.cuda.current_stream(self.env.device))
        with torch.cuda.stream(side), torch.no_grad():      # warm-up outside the graph (allocator, lazy init)
            saved = self.env.snapshot(), self.env._chain.clone(), self.env._stamp.clone(), self.states.clone()
            self.base.copy_(torch.tensor([self.env.steps, self.env._seq, self.env.steps_limit + 1], dtype=torch.int32))
            self._body()
            side.synchronize()
            self.env.restore(saved[0])
            self.env._chain.copy_(saved[1]); self.env._stamp.copy_(saved[2]); self.states.copy_(saved[3])
            if self.on_reset is not None:
                self.on_reset()
            side.synchronize()
        torch.cuda.current_stream(self.env.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._body()
        return self

    def run(self, states, max_replays=None):
        """Play the episode from the observation `states` (of `reset` or the last `step`) to its end.
        Returns (states, rewards) of the LAST REPLAYED step - up to k-1 steps after the one that ended the episode
        (dead and finished cars are frozen there: zero readings, the -0.01 / 0 step rewards), not the terminating
        step's own +1 / -1 rewards; callers that need per-step outputs use `Race.rollout(keep_all=True)` or `step`."""
        env = self.env
        if env._epoch != self.epoch:
            raise GlgError('the environment was reset after this GraphedRollout was created: the graph is bound to '
                           'the previous episode\'s buffers - create a new one')
        if self.graph is None:
            self.capture()
        self.states.copy_(states)
        if self.on_reset is not None:
            self.on_reset()
        start_steps, start_seq = env.steps, env._seq
        self.base.copy_(torch.tensor([start_steps, start_seq, env.steps_limit + 1], dtype=torch.int32))
        replays = 0
        while not env.finished() and (max_replays is None or replays < max_replays):
            self.graph.replay()
            replays += 1
            env.steps += self.k
            env._seq += self.k
            env._alive_known = None
        ran = replays * self.k
        if replays and env.finished():
            # the reference's loop stops right after the step that ended the episode
            env._stamp_host.copy_(env._stamp)
            torch.cuda.current_stream(env.device).synchronize()
            last_alive = int(env._stamp_host.max())                # launch number after which somebody was still alive
            by_death = (max(last_alive, start_seq) - start_seq) + 1 if last_alive >= start_seq else 0
            by_time = env.steps_limit + 1 - start_steps
            env.steps = start_steps + min(ran, by_death, by_time)
            # launches past the time limit were no-ops and left no stamp: the launch counter goes back to the last one
            # that executed, so that "anybody alive" (stamp == launch counter) keeps meaning what it says
            executed = max(0, min(ran, by_time))
            env._seq = start_seq + executed
            env._alive_known = bool((env._stamp_np == env._seq).any()) if executed else None
        if self.hist is not None and env.steps > start_steps:
            env._hist_steps.extend(range(start_steps + 1, env.steps + 1))
        return self.states, self.rewards


class CaptureSafePolicy(torch.nn.Module):
    """Makes a recurrent policy network of the reference (policies/LSTMPolicy.py, ConvLSTMPolicy.py: `state` = list of
    (h, c) tensors that `forward` REBINDS to fresh tensors on every call) usable inside a captured CUDA graph: the
    recurrent state lives in fixed buffers, each forward starts from them and copies the new state back in place.
    Same numbers as the wrapped network; `reset_state`, `state_dict` and `load_state_dict` are the wrapped network's,
    which is all agents/PPOAgent.py:40-63 touches."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self._buf = None

    def reset_state(self):
        if self._buf is None:
            self.net.reset_state()
        else:
            for h, c in self._buf:
                h.zero_()
                c.zero_()

    def forward(self, inputs):
        if self._buf is not None:
            self.net.state = [(h, c) for h, c in self._buf]
        out = self.net(inputs)
        if self._buf is None:                      # first call (eager warm-up): adopt the shapes the network chose
            self._buf = [(h.clone(), c.clone()) for h, c in self.net.state]
        else:
            for (bh, bc), (h, c) in zip(self._buf, self.net.state):
                bh.copy_(h)
                bc.copy_(c)
        self.net.state = None                      # nothing outside the buffers survives the call
        return out

    def state_dict(self, *args, **kwargs):
        return self.net.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.net.load_state_dict(*args, **kwargs)


def capture_safe_agents(agents, deterministic=False):
    """The reference's agents (agents/PPOAgent.py) for `GraphedRollout`: wraps `agent.network` / `agent.old_network` in
    `CaptureSafePolicy` (in place) and returns (act, on_reset) where act(states [P,B,O+2]) -> actions [P,B] is
    train-gan.py:92 - `torch.stack([a.act(s, training=False) for a, s in zip(agents, states)])` - and on_reset() is
    `for a in agents: a.reset()` (train-gan.py:95-96)."""
    for a in agents:
        for name in ('network', 'old_network'):
            net = getattr(a, name, None)
            if net is not None and not isinstance(net, CaptureSafePolicy):
                setattr(a, name, CaptureSafePolicy(net))

    def act(states):
        return torch.stack([a.act(s, deterministic=deterministic, training=False) for a, s in zip(agents, states)], dim=0)

    def on_reset():
        for a in agents:
            a.reset()

    return act, on_reset


class HostStepper(object):
    """`Race.step` for callers whose policies live on the HOST: actions come from and observations go to
    (pinned) host memory every step.  One CUDA graph holds the whole round trip - H2D of the actions and of the
    step counters, the step kernel, D2H of observations, rewards and the alive stamp - so a step costs one
    graph launch and one synchronisation instead of five enqueues.  Same results as `Race.step`, including the
    "nobody alive" early-out (games/race.py:353-356); the returned tensors are views of pinned buffers that the
    next call overwrites.
    """

    def __init__(self, env):
        if env.num_tracks is None or env.num_tracks == 0:
            raise GlgError('HostStepper needs a reset environment with at least one track')
        self.env = env
        self.epoch = env._epoch
        dev = env.device
        B, P, O = env.num_tracks, env.num_players, env.observation_size
        # one input block {step counters (16 B) | actions} and one output block {observations | rewards | alive stamp}
        # on each side, so that the graph is H2D -> step kernel -> D2H
        n_obs, n_rw = P * B * (O + 2), P * B
        self.in_h = torch.zeros((2 + P * B,), dtype=torch.int64).pin_memory()
        self.in_d = torch.zeros((2 + P * B,), dtype=torch.int64, device=dev)
        self.out_h = torch.zeros((n_obs + n_rw + _lib.ALIVE_SLOTS,), dtype=torch.float32).pin_memory()
        self.out_d = torch.zeros((n_obs + n_rw + _lib.ALIVE_SLOTS,), dtype=torch.float32, device=dev)
        self.actions_h = self.in_h[2:].view(P, B)
        self._actions_np = self.actions_h.numpy()
        self._base_np = self.in_h[:2].view(torch.int32).numpy()         # {step number, launch number, last step, -}
        self._base_np[2] = 2 ** 31 - 1                                  # no step limit, like Race.step
        self.states_h = self.out_h[:n_obs].view(P, B, O + 2)
        self.rewards_h = self.out_h[n_obs:n_obs + n_rw].view(P, B)
        self._stamp_np = self.out_h[n_obs + n_rw:].view(torch.int32).numpy()
        base_d = self.in_d[:2].view(torch.int32)
        actions_d = self.in_d[2:].view(P, B)
        states_d = self.out_d[:n_obs].view(P, B, O + 2)
        rewards_d = self.out_d[n_obs:n_obs + n_rw].view(P, B)
        stamp_d = self.out_d[n_obs + n_rw:].view(torch.int32)            # private stamp (launch numbers only grow)
        self.hist = env._history_ring(env.steps_limit + 2)              # recorded board, games/race.py:492-494
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.graph(self.graph, stream=self.stream):
            self.in_d.copy_(self.in_h, non_blocking=True)
            env.step_into(actions_d, states_d, rewards_d, base_d, 1, stamp=stamp_d, history=self.hist)
            self.out_h.copy_(self.out_d, non_blocking=True)

    def step(self, actions):
        """actions: [P,B] integer CPU tensor (or numpy array) -> (states [P,B,O+2], rewards [P,B]) on the host."""
        env = self.env
        if env._epoch != self.epoch:
            raise GlgError('the environment was reset after this HostStepper was created: the graph is bound to the '
                           'previous episode\'s buffers - call host_stepper() again')
        if torch.is_tensor(actions):
            actions = actions.numpy()                      # (CPU tensors only; shares memory)
        if tuple(actions.shape) != tuple(self.actions_h.shape):
            raise ValueError('actions must have shape [num_players, num_boards] = %s' % (tuple(self.actions_h.shape),))
        anybody_alive = env._any_alive()
        if not anybody_alive:                              # games/race.py:353-356 (19-wide quirk), on the host
            env.steps += 1
            P, B, O = env.num_players, env.num_tracks, env.observation_size
            rewards = (1. - env.finishes.float().cpu()) * env.negative_reward
            return torch.zeros((P, B, O + 1), dtype=torch.float32), rewards.t()
        np.copyto(self._actions_np, actions, casting='unsafe')
        self._base_np[0], self._base_np[1] = env.steps, env._seq
        self.graph.replay()                                # on the caller's stream: ordered after its reset / restore
        torch.cuda.current_stream(env.device).synchronize()
        env.steps += 1
        env._seq += 1
        env._alive_known = bool((self._stamp_np == env._seq).any())
        if self.hist is not None and env.steps < self.hist.size(0):
            env._hist_steps.append(env.steps)
        return self.states_h, self.rewards_h


class HostRollout(object):
    """`Race.rollout` for callers whose action tape and observation buffers live on the HOST (pinned memory): T steps in
    chunks of `chunk` steps, pipelined over three streams -

        copy-in  : H2D of chunk c+1's actions
        compute  : the fused rollout kernel of chunk c (one launch per chunk, glg_race_rollout)
        copy-out : D2H of chunk c-1's observations and rewards

    so that the PCIe transfers of neighbouring chunks hide behind the kernel and behind each other (H2D and D2H use the
    two DMA directions).  Every step's actions cross the bus host -> device and every step's observations and rewards
    come back - the same bytes as T calls of `HostStepper.step` - but the host only synchronises once per call.
    Open loop by construction (the tape is given up front); closed-loop host policies use `HostStepper`.
    Results are those of `Race.rollout(actions, keep_all=True)`; bound to one episode like `RolloutPlan`.
    """

    def __init__(self, env, T, chunk=25, mode='fused', first_chunk=None):
        if env.num_tracks is None or env.num_tracks == 0:
            raise GlgError('HostRollout needs a reset environment with at least one track')
        if T <= 0 or chunk <= 0:
            raise ValueError('T and chunk must be positive')
        self.env, self.epoch, self.T = env, env._epoch, int(T)
        dev = env.device
        B, P, O = env.num_tracks, env.num_players, env.observation_size
        self.actions_h = torch.zeros((T, P, B), dtype=torch.int64).pin_memory()
        self.states_h = torch.zeros((T, P, B, O + 2), dtype=torch.float32).pin_memory()
        self.rewards_h = torch.zeros((T, P, B), dtype=torch.float32).pin_memory()
        self.actions_d = torch.zeros((T, P, B), dtype=torch.int64, device=dev)
        self.states_d = torch.empty((T, P, B, O + 2), dtype=torch.float32, device=dev)
        self.rewards_d = torch.empty((T, P, B), dtype=torch.float32, device=dev)
        # The copy-out of the observations is the slowest stage (721 KB per step of config 2 against ~8 us of kernel), so
        # it should start early and then never wait: a short first chunk, full chunks after it.
        first = max(1, chunk // 5) if first_chunk is None else max(1, int(first_chunk))
        cuts = [0] + list(range(min(first, T), T, chunk)) + [T]
        self.bounds = [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]
        self.plans = [env.rollout_plan(self.actions_d[lo:hi], keep_all=True, mode=mode,
                                       out=(self.states_d[lo:hi], self.rewards_d[lo:hi])) for lo, hi in self.bounds]
        self.launches = sum(p.launches for p in self.plans)
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.ev_in = [torch.cuda.Event() for _ in self.bounds]
        self.ev_done = [torch.cuda.Event() for _ in self.bounds]
        self.h2d_bytes = self.actions_h.numel() * 8
        self.d2h_bytes = (self.states_h.numel() + self.rewards_h.numel()) * 4
        self._pending = False

    def submit(self, actions=None):
        """Enqueue one rollout (copies in, kernels, copies out) and return without waiting; `wait()` returns the
        results.  `actions`: [T,P,B] integer CPU tensor / numpy array - a pinned, contiguous int64 tensor is read by the
        copy engine where it lies (it must stay unchanged until `wait()`), anything else goes through this object's pinned
        staging buffer; None = the caller filled `self.actions_h`.  Two `HostRollout`s of one environment used alternately overlap the host's work on call i+1 (and its
        reading of call i's results) with the device's work on call i - each owns its staging buffers and streams; the
        kernels of all calls run in order on the caller's stream, so the episode advances exactly as with `run`."""
        env = self.env
        if env._epoch != self.epoch:
            raise GlgError('the environment was reset after this HostRollout was created - make a new one')
        if self._pending:
            raise GlgError('HostRollout.submit: the previous call has not been waited for')
        src = self.actions_h
        if actions is not None:
            if tuple(actions.shape) != tuple(self.actions_h.shape):
                raise ValueError('actions must have shape [T, num_players, num_boards] = %s' % (tuple(self.actions_h.shape),))
            if (torch.is_tensor(actions) and actions.dtype == torch.int64 and actions.is_contiguous()
                    and not actions.is_cuda and actions.is_pinned()):
                src = actions                          # already pinned int64: copied to the device straight from it
            else:
                a = actions.numpy() if torch.is_tensor(actions) else np.asarray(actions)
                np.copyto(self.actions_h.numpy(), a, casting='unsafe')
        main = torch.cuda.current_stream(env.device)
        with torch.no_grad():
            for c, (lo, hi) in enumerate(self.bounds):
                with torch.cuda.stream(self.s_in):
                    self.actions_d[lo:hi].copy_(src[lo:hi], non_blocking=True)
                    self.ev_in[c].record(self.s_in)
            for c, (lo, hi) in enumerate(self.bounds):
                main.wait_event(self.ev_in[c])
                self.plans[c].run()                        # on the caller's stream: ordered after its reset / restore
                self.ev_done[c].record(main)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.ev_done[c])
                    self.states_h[lo:hi].copy_(self.states_d[lo:hi], non_blocking=True)
                    self.rewards_h[lo:hi].copy_(self.rewards_d[lo:hi], non_blocking=True)
        self._pending = True
        return self

    def wait(self):
        """-> (states [T,P,B,O+2], rewards [T,P,B]) of the submitted call: pinned host tensors, complete on return,
        overwritten by this object's next call."""
        if not self._pending:
            raise GlgError('HostRollout.wait: nothing was submitted')
        self.s_out.synchronize()                           # the copies out are the last thing a call does
        self._pending = False
        return self.states_h, self.rewards_h

    def run(self, actions=None):
        """actions: [T,P,B] integer CPU tensor / numpy array (copied into the pinned staging buffer), or None when the
        caller filled `self.actions_h` itself.  -> (states [T,P,B,O+2], rewards [T,P,B]): pinned host tensors, complete
        when the call returns, overwritten by the next call.  (`submit` + `wait`.)"""
        return self.submit(actions).wait()
