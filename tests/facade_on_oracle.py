"""Test scaffolding (not a test module): runs the drop-in ``Environment`` facade on a CPU box by replacing the native
context with one that computes every launch with the oracle.  Only the *Python* side of the facade is under test here
(attribute packing, exception mapping, word-offset and look-ahead bookkeeping, np_game painting); the kernels are tested
on the GPU.  Used by tests/test_reference_agents_cpu.py in a subprocess, where the reference's own agents
(/root/reference/agents, unmodified) import ``shipping`` from compat/ and play on this facade."""
import ctypes as C

import numpy as np
import torch

from oracle import lib as orc
import shippingenv_b200.environment as E
from shippingenv_b200 import _native as nat


def _view(ptr, dtype, count=1):
    dt = np.dtype(dtype)
    buf = (C.c_uint8 * (dt.itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=count)


class OracleContext:
    """The slice of ``_native.Context`` the facade uses, computed by the oracle on host memory."""

    def __init__(self, device):
        self.device, self.P, self.o, self._map, self._ports = device, -1, None, None, None
        self.launches = 0

    def set_map(self, passable):
        self._map, self.o = np.array(passable, dtype=np.uint8), None

    def set_ports(self, xy, fuel, cargo):
        self._ports, self.P, self.o = (list(xy), list(fuel), list(cargo)), len(xy), None

    def _oracle(self):
        if self.o is None:
            self.o = orc.Oracle(self._map, *self._ports)
        return self.o

    def reset(self, ships, clock, seed, env_id_base, n, mask_bits, stream):
        self.launches += 1
        self._oracle().reset(_view(ships, orc.SHIP_DTYPE, n), _view(clock, np.uint32, n), seed, env_id_base)

    def sample_action(self, ships, clock, actions, words_used, status, seed, env_id_base, n, word0, stream):
        self.launches += 1
        a, wu, st = self._oracle().sample_action(_view(ships, orc.SHIP_DTYPE, n), _view(clock, np.uint32, n), seed,
                                                 env_id_base, word0)
        _view(actions, np.int16, n)[:] = a
        _view(words_used, np.uint8, n)[:] = wu
        _view(status, np.uint8, n)[:] = st

    def step_sample(self, ships, clock, actions, io, next_action, next_words, next_status, seed, env_id_base, n,
                    max_steps, flags, word0, stream):
        self.launches += 1
        o = self._oracle()
        s, c = _view(ships, orc.SHIP_DTYPE, n), _view(clock, np.uint32, n)
        r = o.step(s, c, _view(actions, np.int16, n).copy(), seed, env_id_base, max_steps, bool(flags & 1), word0)
        _view(io.reward_f64, np.float64, n)[:] = r["reward"]
        _view(io.status, np.uint8, n)[:] = r["status"]
        _view(io.done_bits, np.uint32)[0] = int(r["done"][0])
        _view(io.reset_bits, np.uint32)[0] = int(r["reset"][0])
        _view(io.error_bits, np.uint32)[0] = int(r["status"][0] != 0)
        a, wu, st = o.sample_action(s, c, seed, env_id_base, word0 if r["status"][0] != 0 else 0)
        _view(next_action, np.int16, n)[:] = a
        _view(next_words, np.uint8, n)[:] = wu
        _view(next_status, np.uint8, n)[:] = st

    def serve_step(self, block, seed, env_id, max_steps, flags, word0):
        """shipenv_serve_step on the 128-byte host block (include/shipenv.h: shipenv_host_block)."""
        self.launches += 1
        o = self._oracle()
        s, c = _view(block, orc.SHIP_DTYPE), _view(block + 32, np.uint32)
        r = o.step(s, c, _view(block + 48, np.int16).copy(), seed, env_id, max_steps, bool(flags & 1), word0)
        _view(block + 16, np.float64)[:] = r["reward"]
        _view(block + 76, np.uint8)[:] = r["status"]
        _view(block + 64, np.uint32)[0] = int(r["done"][0])
        _view(block + 68, np.uint32)[0] = int(r["reset"][0])
        _view(block + 72, np.uint32)[0] = int(r["status"][0] != 0)
        a, wu, st = o.sample_action(s, c, seed, env_id, word0 if r["status"][0] != 0 else 0)
        _view(block + 84, np.int16)[:] = a
        _view(block + 86, np.uint8)[:] = wu
        _view(block + 87, np.uint8)[:] = st

    def synchronize(self, stream):
        pass


class _Stream:
    cuda_stream = 0


class _Pinnable:
    def __init__(self, t):
        self.t = t

    def pin_memory(self):
        return self.t


class _Cuda:
    is_available = staticmethod(lambda: True)
    current_device = staticmethod(lambda: 0)
    current_stream = staticmethod(lambda device=None: _Stream())


class _TorchForFacade:
    """What shippingenv_b200.environment sees as ``torch``: real torch, except that a CUDA device seems present and
    pinning is a no-op.  The rest of the process (e.g. the reference's DQN agent) keeps the real module."""
    cuda = _Cuda()

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def zeros(*a, **k):
        return _Pinnable(torch.zeros(*a, **k))


_keep = []          # ctypes callbacks must outlive the C code that holds their address


def install(fast=None):
    """Patch the facade module so that it runs without a GPU.  Returns the dict of fake contexts (per device).
    ``fast``: True = the facade's C fast path (csrc/facade_fast.c) is bound to a callback that lets the oracle answer
    ``shipenv_serve_step`` (so the C code itself is exercised on a CPU box), False = Python only, None = the
    environment decides (SHIPENV_NO_FASTPATH)."""
    contexts = {}
    by_handle = {}

    def context(device_index):
        if device_index not in contexts:
            ctx = OracleContext(device_index)
            ctx._h_int = 0x1000 + device_index            # what the C fast path passes back as the context handle
            by_handle[ctx._h_int] = ctx
            contexts[device_index] = [ctx, None]
        return contexts[device_index]

    E._context = context
    E.torch = _TorchForFacade()
    E.nat.host_device_pointer = lambda host_ptr: host_ptr
    import os
    if fast is None:
        fast = os.environ.get("SHIPENV_NO_FASTPATH") != "1"
    E._fast, E._fast_tried = None, True                   # never the real shipenv_serve_step on this box
    if fast:
        from shippingenv_b200 import _facade_fast as mod

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32)
        def serve(handle, block, seed, env_id, max_steps, flags, word0):
            by_handle[handle].serve_step(block, seed, env_id, max_steps, flags, word0)
            return 0

        _keep.append(serve)
        mod.bind(C.cast(serve, C.c_void_p).value, E._ERRORS, nat.check, E.MOVES, nat.N_CARGO_ACTIONS,
                 nat.N_FUEL_ACTIONS, nat.ACT_EXT_CARGO, nat.ACT_EXT_FUEL)
        E._fast = mod
    return contexts
