"""``BatchedEnvironment``: N independent ShippingEnv ships resident in HBM, advanced by the
sm_100a kernels behind the C ABI (include/shipenv.h).  PyTorch only owns the memory and the
stream; every transition is computed by ``libshipenv.so``.

Semantics are the reference's ``Environment`` (shipping/environment.py) per env, plus the
batched contract of DESIGN.md: a per-env clock (episode << 16 | step), status codes instead of
exceptions, optional auto-reset, policy actions (-1 = ``sample_action`` until accepted).
"""
from dataclasses import dataclass
from typing import Optional

import os

import numpy as np
import torch

from . import _native as nat
from .mapio import derive_grid, passable_of
from .stats import new_stats, reduce_stats, summarize

# reference defaults: utils/constants.py:57-63 and the seed-2024 stock draw
DEFAULT_PORTS = [[41, 40], [60, 22], [78, 29], [49, 72], [62, 72]]
DEFAULT_PORT_FUEL = [20, 14, 18, 12, 16]
DEFAULT_PORT_CARGO = [10, 11, 13, 20, 18]


@dataclass
class StepResult:
    reward: torch.Tensor                 # float32 [N]
    done_bits: torch.Tensor              # int32 [ceil(N/32)], bit i%32 of word i//32
    reset_bits: torch.Tensor
    error_bits: torch.Tensor
    reward64: Optional[torch.Tensor] = None
    status: Optional[torch.Tensor] = None
    taken: Optional[torch.Tensor] = None
    obs: Optional[torch.Tensor] = None


def unpack_bits(bits, n):
    """int32 ballot words -> bool [n] (on the words' device)."""
    shifts = torch.arange(32, device=bits.device, dtype=torch.int32)
    return ((bits[:, None] >> shifts) & 1).reshape(-1)[:n].bool()


def _ptr(t):
    return None if t is None else t.data_ptr()


class BatchedEnvironment:
    def __init__(self, num_envs, passable=None, map_path=None, game_size=(100, 100), ports=None,
                 port_fuel=None, port_cargo=None, seed=0x5EED, device=None, env_id_base=0,
                 max_steps=1000, auto_reset=True, consts=None, track_stats=False):
        if not torch.cuda.is_available():
            raise nat.ShipEnvError("BatchedEnvironment needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n = int(num_envs)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.env_id_base = int(env_id_base)
        self.max_steps = int(max_steps)
        self.auto_reset = bool(auto_reset)
        if passable is None:
            if map_path is None:
                raise ValueError("give either `passable` (uint8 [W,H]) or `map_path`")
            passable = passable_of(derive_grid(map_path, game_size))
        self.passable = np.ascontiguousarray(passable, dtype=np.uint8)
        ports = DEFAULT_PORTS if ports is None else ports
        P = len(ports)
        if port_fuel is None or port_cargo is None:
            if ports == DEFAULT_PORTS and port_fuel is None and port_cargo is None:
                port_fuel, port_cargo = DEFAULT_PORT_FUEL, DEFAULT_PORT_CARGO
            else:  # add_port draws, shipping/environment.py:63-64
                stock = [nat.port_stock(self.seed, k) for k in range(P)]
                port_fuel = [s[0] for s in stock] if port_fuel is None else port_fuel
                port_cargo = [s[1] for s in stock] if port_cargo is None else port_cargo
        W, H = self.passable.shape
        for x, y in ports:
            if not (0 <= x < W and 0 <= y < H):
                raise ValueError("Coordinates not within map")           # shipping/environment.py:60
        self.ports = [list(map(int, p)) for p in ports]
        self.port_fuel, self.port_cargo = list(map(int, port_fuel)), list(map(int, port_cargo))
        self.ctx = nat.Context(self.device.index)
        self.ctx.set_map(self.passable)
        self.ctx.set_ports(self.ports, self.port_fuel, self.port_cargo)
        if consts:
            self.ctx.set_consts(**consts)
        self.P = P
        self.obs_width = 6 + 4 * P
        self.num_actions = self.ctx.num_actions

        dev, n, nw = self.device, self.n, (self.n + 31) // 32
        self.ships = torch.zeros((n, 16), dtype=torch.uint8, device=dev)      # shipenv_ship records
        self.clock = torch.zeros(n, dtype=torch.int32, device=dev)             # u32 bit pattern
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.done_bits = torch.zeros(nw, dtype=torch.int32, device=dev)
        self.reset_bits = torch.zeros(nw, dtype=torch.int32, device=dev)
        self.error_bits = torch.zeros(nw, dtype=torch.int32, device=dev)
        self._reward64 = self._status = self._taken = self._obs = self._valid = self._keys = None
        self.ep_return = self.stats = None
        if track_stats:
            self.ep_return = torch.zeros(n, dtype=torch.float64, device=dev)
            self.stats = new_stats(dev)
        self._policy = None
        # chained launches (include/shipenv.h, shipenv_step_io.chain): per-CTA launch counters of this env group
        self._chain = None
        if os.environ.get("SHIPENV_NO_CHAIN") != "1":
            self._chain = torch.zeros(nat.lib().shipenv_chain_words(n), dtype=torch.int32, device=dev)
        self._chain_ok = False             # cleared whenever the state tensors are written outside step()
        self._host_out = self._host_ws = None
        self.stream, self._stream_handle = None, None
        self._step_cache, self._io_keepalive = {}, []
        self._lib_step = nat.lib().shipenv_step
        self._ctx_handle = self.ctx._h
        self._ships_ptr, self._clock_ptr = self.ships.data_ptr(), self.clock.data_ptr()

    # ---- views of the ship records ------------------------------------------------------
    @property
    def x(self): return self.ships[:, 0]
    @property
    def y(self): return self.ships[:, 1]
    @property
    def origin(self): return self.ships[:, 2]
    @property
    def dest(self): return self.ships[:, 3]
    @property
    def cargo(self): return self.ships.view(torch.int32)[:, 1]
    @property
    def fuel(self): return self.ships.view(torch.float64)[:, 1]

    def break_chain(self):
        """Call after writing ``ships`` / ``clock`` with anything but this object's methods: the next step then waits
        for all earlier work of its stream even if it is ``chained``."""
        self._chain_ok = False

    def use_stream(self, stream):
        """Launch everything of this environment on ``stream`` (a torch.cuda.Stream) instead of the current stream;
        ``None`` goes back to the current stream.  Independent env groups on separate streams overlap their launch
        boundaries; the caller orders them against producers / consumers of their tensors (wait_stream / events)."""
        self._chain_ok = False
        self.stream = stream
        self._stream_handle = None if stream is None else stream.cuda_stream

    def _stream(self):
        h = self._stream_handle
        return h if h is not None else torch.cuda.current_stream(self.device).cuda_stream

    def _buf(self, name, shape, dtype):
        t = getattr(self, name)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            setattr(self, name, t)
        return t

    # ---- the reference API, batched -----------------------------------------------------
    def reset(self, mask_bits=None):
        """``Environment.reset`` for all envs (or those whose mask bit is set)."""
        self._chain_ok = False
        self.ctx.reset(self.ships.data_ptr(), self.clock.data_ptr(), self.seed, self.env_id_base, self.n,
                       _ptr(mask_bits), self._stream())
        if self.ep_return is not None and mask_bits is None:
            self.ep_return.zero_()

    def step(self, actions, want_obs=False, want_status=False, want_taken=False, want_reward64=False,
             obs_out=None, word0=0, chained=False):
        """``Environment.step`` for every env.  ``actions``: int16 [N] flat indices on this device.

        ``chained=True`` lets this launch overlap the previous ``step`` of this environment at the launch boundary
        (``SHIPENV_F_CHAINED``): CTA b starts as soon as CTA b of the previous chained step has finished its share of
        the envs, instead of after the whole previous grid.  Only for ``actions`` that were complete before the previous step was enqueued
        (pre-generated action streams); policy steps (``step_policy``) are always chained, they read no actions.
        Results are identical either way."""
        if actions.dtype != torch.int16 or actions.device != self.device or actions.numel() != self.n:
            raise ValueError("actions must be an int16 tensor of shape [num_envs] on the env's device")
        if not actions.is_contiguous():
            actions = actions.contiguous()
        # the chain array travels only with launches that ask for it: every other step keeps the ticket-scheduled grid
        use_chain = (chained or actions is self._policy) and self._chain is not None
        key = (want_obs, want_status, want_taken, want_reward64, None if obs_out is None else obs_out.data_ptr(),
               use_chain)
        cached = self._step_cache.get(key)
        if cached is None:
            if len(self._step_cache) >= 64:        # a caller cycling through many obs_out buffers: start over
                self._step_cache.clear()
                self._io_keepalive.clear()
            cached = self._step_cache[key] = self._build_step_io(want_obs, want_status, want_taken, want_reward64,
                                                                 obs_out, use_chain)
        io_ref, res = cached
        flags = (nat.F_AUTO_RESET if self.auto_reset else 0) | (nat.F_ALL_POLICY if actions is self._policy else 0)
        if use_chain and self._chain_ok:
            flags |= nat.F_CHAINED         # the library honours it only right behind another chained step of this env
        self._chain_ok = use_chain
        rc = self._lib_step(self._ctx_handle, self._ships_ptr, self._clock_ptr, actions.data_ptr(), io_ref, self.seed,
                            self.env_id_base, self.n, self.max_steps, flags,
                            word0, self._stream())
        if rc:
            nat.check(rc)
        return res

    def _build_step_io(self, want_obs, want_status, want_taken, want_reward64, obs_out, use_chain=False):
        import ctypes
        io = nat.StepIO()
        io.reward_f32 = self.reward.data_ptr()
        io.done_bits, io.reset_bits, io.error_bits = (self.done_bits.data_ptr(), self.reset_bits.data_ptr(),
                                                      self.error_bits.data_ptr())
        res = StepResult(self.reward, self.done_bits, self.reset_bits, self.error_bits)
        if want_reward64:
            res.reward64 = self._buf("_reward64", self.n, torch.float64)
            io.reward_f64 = res.reward64.data_ptr()
        if want_status:
            res.status = self._buf("_status", self.n, torch.uint8)
            io.status = res.status.data_ptr()
        if want_taken:
            res.taken = self._buf("_taken", self.n, torch.int16)
            io.taken = res.taken.data_ptr()
        if want_obs or obs_out is not None:
            res.obs = obs_out if obs_out is not None else self._buf("_obs", (self.n, self.obs_width), torch.float32)
            if res.obs.dtype != torch.float32 or tuple(res.obs.shape) != (self.n, self.obs_width) \
                    or not res.obs.is_contiguous():
                raise ValueError("obs_out must be a contiguous float32 [num_envs, 6+4P] tensor")
            io.obs = res.obs.data_ptr()
        if self.ep_return is not None:
            io.ep_return, io.stats = self.ep_return.data_ptr(), self.stats.data_ptr()
        if use_chain:
            io.chain = self._chain.data_ptr()
        self._io_keepalive.append(io)
        return ctypes.byref(io), res

    def host_buffers(self):
        """Page-locked host outputs for :meth:`step_host`, carved from one allocation in the order ``reward`` float32 [N],
        ``done_bits``, ``reset_bits``, ``error_bits`` int32 [ceil(N/32)] -- laid out like this they come back with a
        single device-to-host copy."""
        nw = (self.n + 31) // 32
        block = torch.empty(4 * self.n + 12 * nw, dtype=torch.uint8).pin_memory()
        reward = block[:4 * self.n].view(torch.float32)
        masks = block[4 * self.n:].view(torch.int32)
        return reward, masks[:nw], masks[nw:2 * nw], masks[2 * nw:]

    def step_host(self, actions, reward=None, done_bits=None, reset_bits=None, error_bits=None, chunks=0):
        """``Environment.step`` for every env with the actions in HOST memory and the results delivered to HOST memory
        (``shipenv_step_host``): ``actions`` is an int16 CPU tensor [N]; ``reward`` float32 [N] and the three int32
        ballot arrays [ceil(N/32)] are CPU tensors that are filled in place (allocated page-locked on first use when
        not given).  Copies and kernels of successive chunks of the batch are pipelined inside the call
        (``chunks`` = 0: chosen from the batch size; 1: no chunking, for callers that already overlap several
        environments on their own streams); it is asynchronous with respect to the host when the buffers are
        page-locked -- synchronise the stream before reading them."""
        if actions.dtype != torch.int16 or actions.device.type != "cpu" or actions.numel() != self.n \
                or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous int16 CPU tensor of shape [num_envs]")
        nw = (self.n + 31) // 32
        if self._host_out is None:
            self._host_out = self.host_buffers()
        if self._host_ws is None:
            size = nat.lib().shipenv_step_host_workspace(self.n)
            self._host_ws = torch.empty(size, dtype=torch.uint8, device=self.device)
        outs = [given if given is not None else own
                for given, own in zip((reward, done_bits, reset_bits, error_bits), self._host_out)]
        for t, (dt, k) in zip(outs, ((torch.float32, self.n), (torch.int32, nw), (torch.int32, nw), (torch.int32, nw))):
            if t.dtype != dt or t.device.type != "cpu" or t.numel() != k or not t.is_contiguous():
                raise ValueError("host outputs must be contiguous CPU tensors: float32 [N], int32 [ceil(N/32)] x 3")
        flags = nat.F_AUTO_RESET if self.auto_reset else 0
        self.ctx.step_host(self._ships_ptr, self._clock_ptr, actions.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                           outs[2].data_ptr(), outs[3].data_ptr(), self._host_ws.data_ptr(), self._host_ws.numel(),
                           self.seed, self.env_id_base, self.n, self.max_steps, flags, int(chunks), self._stream())
        return StepResult(*outs)

    def capture_steps(self, actions, **kw):
        """Capture ``len(actions)`` consecutive steps (``actions``: int16 [T, N], one row per step) into a CUDA
        graph and return it; every ``graph.replay()`` advances all envs by T steps with one launch from the host.
        The step launches hold no host-side state (the tile scheduler resets itself on the device), so a captured
        sequence can be replayed any number of times.  For small batches (BASELINE configs[1]: 4,096 envs) the
        step is launch-bound and the graph removes the per-step host cost."""
        if actions.dim() != 2 or actions.shape[1] != self.n or actions.dtype != torch.int16:
            raise ValueError("actions must be an int16 tensor of shape [T, num_envs]")
        rows = [actions[t] for t in range(actions.shape[0])]
        self._chain_ok = False
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for row in rows:                      # the whole [T, N] block exists already: steps 1.. may chain
                self.step(row, chained=True, **kw)
        graph._shipenv_keepalive = (actions, rows)
        self._chain_ok = False
        return graph

    def step_many(self, actions=None, steps=None, want_status=False, want_taken=False, want_reward64=False,
                  want_masks=True):
        """``steps`` consecutive ``Environment.step`` calls per env in ONE launch (``shipenv_step_many``): the ships stay
        in registers between steps.  ``actions``: int16 [T, N] on this device, or ``None`` for T = ``steps`` steps of
        the random policy.  Same results as T calls of :meth:`step`.  Returns a dict of [T, ...] tensors: ``reward``
        float32 [T, N], ``done_bits`` / ``reset_bits`` / ``error_bits`` int32 [T, ceil(N/32)], optionally ``status``,
        ``taken``, ``reward64``."""
        if actions is not None:
            if actions.dim() != 2 or actions.shape[1] != self.n or actions.dtype != torch.int16 \
                    or actions.device != self.device:
                raise ValueError("actions must be an int16 tensor of shape [T, num_envs] on the env's device")
            actions = actions.contiguous()
            T = actions.shape[0]
        else:
            if steps is None:
                raise ValueError("give `actions` [T, N] or `steps` (random policy)")
            T = int(steps)
        dev, n, nw = self.device, self.n, (self.n + 31) // 32
        out = dict(reward=torch.empty((T, n), dtype=torch.float32, device=dev))
        if want_masks:
            for k in ("done_bits", "reset_bits", "error_bits"):
                out[k] = torch.empty((T, nw), dtype=torch.int32, device=dev)
        if want_status:
            out["status"] = torch.empty((T, n), dtype=torch.uint8, device=dev)
        if want_taken:
            out["taken"] = torch.empty((T, n), dtype=torch.int16, device=dev)
        if want_reward64:
            out["reward64"] = torch.empty((T, n), dtype=torch.float64, device=dev)
        flags = (nat.F_AUTO_RESET if self.auto_reset else 0) | (nat.F_ALL_POLICY if actions is None else 0)
        self._chain_ok = False
        self.ctx.step_many(self._ships_ptr, self._clock_ptr, _ptr(actions), T, out["reward"].data_ptr(),
                           _ptr(out.get("done_bits")), _ptr(out.get("reset_bits")), _ptr(out.get("error_bits")),
                           _ptr(out.get("status")), _ptr(out.get("taken")), _ptr(out.get("reward64")), self.seed,
                           self.env_id_base, n, self.max_steps, flags, self._stream())
        return out

    def step_policy(self, **kw):
        """One step of the reference's random policy (``sample_action`` + retry) on every env."""
        if self._policy is None:
            self._policy = torch.full((self.n,), nat.ACT_POLICY, dtype=torch.int16, device=self.device)
        return self.step(self._policy, **kw)

    def observe(self, out=None):
        """``preprocess_state(env._build_state())`` rows, float32 [N, 6+4P]."""
        obs = out if out is not None else self._buf("_obs", (self.n, self.obs_width), torch.float32)
        self.ctx.observe(self.ships.data_ptr(), obs.data_ptr(), self.n, self._stream())
        return obs

    def valid_action_bits(self, out=None):
        """``DQNAgent.is_valid_action`` for all actions of all envs: int32 [N, ceil(A/32)] bit words."""
        words = (self.num_actions + 31) // 32
        m = out if out is not None else self._buf("_valid", (self.n, words), torch.int32)
        self.ctx.valid_actions(self.ships.data_ptr(), m.data_ptr(), self.n, self._stream())
        return m

    def valid_action_mask(self):
        """The same as a bool tensor [N, A] (for ``q.masked_fill(~mask, -inf)``)."""
        bits = self.valid_action_bits()
        shifts = torch.arange(32, device=self.device, dtype=torch.int32)
        return ((bits[:, :, None] >> shifts) & 1).reshape(self.n, -1)[:, :self.num_actions].bool()

    def discretize(self, out=None):
        """``discretize_state`` keys of the tabular agents: int16 [N, 6] = x, y, "cargo" bucket, fuel bucket,
        origin | -1, dest | -1 (utils/preprocessing.py:65-90)."""
        k = out if out is not None else self._buf("_keys", (self.n, 6), torch.int16)
        self.ctx.discretize(self.ships.data_ptr(), k.data_ptr(), self.n, self._stream())
        return k

    def sample_action(self, word0=0):
        """``Environment.sample_action`` per env -> (actions int16 [N], words_used uint8 [N], status uint8 [N])."""
        a = torch.empty(self.n, dtype=torch.int16, device=self.device)
        wu = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        st = torch.empty(self.n, dtype=torch.uint8, device=self.device)
        self.ctx.sample_action(self.ships.data_ptr(), self.clock.data_ptr(), a.data_ptr(), wu.data_ptr(),
                               st.data_ptr(), self.seed, self.env_id_base, self.n, word0, self._stream())
        return a, wu, st

    def rollout(self, sims, max_steps=100, roots=None, root_id_base=None, want_final=False):
        """``MCTSAgent._rollout`` x ``sims`` from every root (default: the current ships).

        Returns a dict with ``returns`` float64 [R, sims] and, if asked, ``final`` uint8
        [R, sims, 16], plus ``steps`` int32 and ``status`` uint8 [R, sims]."""
        roots = self.ships if roots is None else roots
        R = roots.shape[0]
        dev = self.device
        ret = torch.empty((R, sims), dtype=torch.float64, device=dev)
        steps = torch.empty((R, sims), dtype=torch.int32, device=dev)
        status = torch.empty((R, sims), dtype=torch.uint8, device=dev)
        final = torch.empty((R, sims, 16), dtype=torch.uint8, device=dev) if want_final else None
        base = self.env_id_base if root_id_base is None else root_id_base
        self.ctx.rollout(roots.data_ptr(), R, sims, max_steps, ret.data_ptr(), _ptr(final), steps.data_ptr(),
                         status.data_ptr(), self.seed, base, self._stream())
        return dict(returns=ret, final=final, steps=steps, status=status)

    def episode_summary(self, group=None):
        """Episode statistics since construction, reduced over the process group (the one collective)."""
        if self.stats is None:
            raise ValueError("construct with track_stats=True")
        return summarize(reduce_stats(self.stats, group))

    def mcts(self, simulations=50, max_rollout_steps=100, c_param=1.4, roots=None, root_id_base=None):
        """``MCTSAgent.choose_action`` for every root (default: the current ships), trees on the device.

        Returns dict: ``action`` int16 [R] (flat move index, -1 without children), ``root_visits`` int32 [R],
        ``child_action`` / ``child_visits`` / ``child_total`` [R, 4] in expansion order, ``status`` uint8 [R]."""
        roots = self.ships if roots is None else roots
        R, dev = roots.shape[0], self.device
        out = dict(action=torch.empty(R, dtype=torch.int16, device=dev),
                   root_visits=torch.empty(R, dtype=torch.int32, device=dev),
                   child_action=torch.empty((R, 4), dtype=torch.int16, device=dev),
                   child_visits=torch.empty((R, 4), dtype=torch.int32, device=dev),
                   child_total=torch.empty((R, 4), dtype=torch.float64, device=dev),
                   status=torch.empty(R, dtype=torch.uint8, device=dev))
        # numpy.log, so that UCB1 is bit-identical to the reference's np.log (agents/mcts.py:125)
        log_table = np.concatenate([[0.0], np.log(np.arange(1, simulations + 2))])
        base = self.env_id_base if root_id_base is None else root_id_base
        self.ctx.mcts(roots.data_ptr(), R, simulations, max_rollout_steps, c_param, log_table.ctypes.data,
                      out["action"].data_ptr(), out["root_visits"].data_ptr(), out["child_action"].data_ptr(),
                      out["child_visits"].data_ptr(), out["child_total"].data_ptr(), out["status"].data_ptr(),
                      self.seed, base, self._stream())
        return out

    # ---- host <-> device state ----------------------------------------------------------
    def load_state(self, ships_np, clock_np=None):
        """Overwrite the ship records (numpy structured/uint8 [N,16]) and optionally the clock."""
        self._chain_ok = False
        raw = np.ascontiguousarray(ships_np).view(np.uint8).reshape(self.n, 16)
        self.ships.copy_(torch.from_numpy(raw))
        if clock_np is not None:
            self.clock.copy_(torch.from_numpy(np.ascontiguousarray(clock_np, dtype=np.uint32).view(np.int32)))

    def state_numpy(self):
        """(ships uint8 [N,16], clock uint32 [N]) on the host."""
        return self.ships.cpu().numpy(), self.clock.cpu().numpy().view(np.uint32)
