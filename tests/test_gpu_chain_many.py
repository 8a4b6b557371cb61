"""GPU: chained launches (shipenv_step_io.chain + SHIPENV_F_CHAINED), the fused T-step entry point
(shipenv_step_many) and the throughput kernel at the benchmark's size, all against the CPU oracle.
Bit-exact for ship records, clocks, f64/f32 rewards, status, taken action and the three masks."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import pytest

from conftest import golden
from oracle import lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


class ThreadedOracle:
    """The C oracle over slices of the env range on all host threads (the step is per env)."""

    def __init__(self, z, ports, fuel, cargo, n, seed, env_base, max_steps):
        self.o = orc.Oracle(z["passable"], ports, fuel, cargo)
        self.n, self.seed, self.base, self.max_steps = n, seed, env_base, max_steps
        self.ships, self.clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
        self.threads = min(os.cpu_count() or 1, 32)
        self.bounds = np.linspace(0, n, self.threads + 1).astype(np.int64)
        self.pool = ThreadPoolExecutor(self.threads)
        self.o.reset(self.ships, self.clock, seed, env_base)

    def step(self, actions):
        n = self.n
        out = dict(reward=np.zeros(n, np.float64), done=np.zeros(n, np.uint8), status=np.zeros(n, np.uint8),
                   reset=np.zeros(n, np.uint8), taken=np.zeros(n, np.int16))
        actions = np.ascontiguousarray(actions, dtype=np.int16)
        L = orc.lib()

        def work(k):
            lo, hi = int(self.bounds[k]), int(self.bounds[k + 1])
            if hi == lo:
                return
            L.orc_step(C.byref(self.o.w), C.c_void_p(self.ships.ctypes.data + 16 * lo),
                       C.c_void_p(self.clock.ctypes.data + 4 * lo), C.c_void_p(actions.ctypes.data + 2 * lo),
                       C.c_void_p(out["reward"].ctypes.data + 8 * lo), C.c_void_p(out["done"].ctypes.data + lo),
                       C.c_void_p(out["status"].ctypes.data + lo), C.c_void_p(out["reset"].ctypes.data + lo),
                       C.c_void_p(out["taken"].ctypes.data + 2 * lo), C.c_uint64(self.seed),
                       C.c_uint64(self.base + lo), C.c_int64(hi - lo), C.c_int(self.max_steps), C.c_int(1), C.c_uint32(0))
        list(self.pool.map(work, range(self.threads)))
        return out


def bits(t, n):
    from shippingenv_b200 import unpack_bits
    return unpack_bits(t, n).cpu().numpy()


def test_bench_kernel_at_bench_size_vs_oracle(torch_cuda, default_map):
    """step_fast_kernel<move actions, no obs, 2 envs per thread, lean> -- the kernel bench.py times -- at
    2,097,152 envs with env_id_base = 3 * 2^21 (rank 3 of the 8-GPU run): every output of every step against the
    multithreaded C oracle.  max_steps = 5 puts auto-resets (the deferred general path) into every step."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    n, seed, base, max_steps = 1 << 21, 0x5EED, 3 << 21, 5
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", env_id_base=base,
                             max_steps=max_steps, auto_reset=True)
    ref = ThreadedOracle(default_map, env.ports, env.port_fuel, env.port_cargo, n, seed, base, max_steps)
    env.reset()
    assert env.state_numpy()[0].tobytes() == ref.ships.tobytes()
    gen = torch.Generator(device="cuda").manual_seed(99)
    n_reset = 0
    for t in range(8):
        a = torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.int16)
        r = env.step(a)                               # reward_f32 + masks only: the lean variant
        want = ref.step(a.cpu().numpy())
        g_ships, g_clock = env.state_numpy()
        assert g_ships.tobytes() == ref.ships.tobytes(), f"ships differ at step {t}"
        assert (g_clock == ref.clock).all()
        assert (r.reward.cpu().numpy() == want["reward"].astype(np.float32)).all()
        assert (bits(r.done_bits, n) == want["done"].astype(bool)).all()
        assert (bits(r.reset_bits, n) == want["reset"].astype(bool)).all()
        assert (bits(r.error_bits, n) == (want["status"] != 0)).all()
        n_reset += int(want["reset"].sum())
    assert n_reset >= n                               # every env was reset at least once


@pytest.mark.parametrize("n,max_steps,policy", [(1 << 20, 7, False), (300_001, 3, False), (4096, 11, False),
                                                (777, 5, False), (1 << 19, 9, True), (70_001, 4, True)])
def test_chained_launches_match_oracle(torch_cuda, default_map, n, max_steps, policy):
    """T steps enqueued back to back with SHIPENV_F_CHAINED (CTA b of each launch starts as soon as CTA b of the previous
    launch has finished its static env range): the final state and the last step's outputs equal the oracle's after T steps.
    Short episodes keep the deferred general path busy in every launch."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    seed, T = 4242, 24
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=max_steps,
                             auto_reset=True)
    assert env._chain is not None
    ref = ThreadedOracle(default_map, env.ports, env.port_fuel, env.port_cargo, n, seed, 0, max_steps)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(5)
    # one tensor per step: a row of a [T, n] block is not 16-byte aligned for odd n (TMA bulk copies need that)
    acts = [torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.int16) for _ in range(T)]
    if policy:
        for a in acts:
            a.fill_(-1)
    torch.cuda.synchronize()
    for t in range(T):                                # nothing between the launches: they overlap on the device
        r = env.step_policy() if policy else env.step(acts[t], chained=True)
    a_np = np.stack([a.cpu().numpy() for a in acts])
    for t in range(T):
        want = ref.step(a_np[t])
    g_ships, g_clock = env.state_numpy()
    assert g_ships.tobytes() == ref.ships.tobytes()
    assert (g_clock == ref.clock).all()
    assert (r.reward.cpu().numpy() == want["reward"].astype(np.float32)).all()
    assert (bits(r.done_bits, n) == want["done"].astype(bool)).all()
    assert (bits(r.reset_bits, n) == want["reset"].astype(bool)).all()
    assert (bits(r.error_bits, n) == (want["status"] != 0)).all()
    # the chain array is quiescent: every CTA of the grid has arrived at and finished T launches
    chain = env._chain.cpu().numpy().view(np.uint32)
    word = int(chain[0]) | int(chain[1]) << 32       # CTAs of all launches so far << 16 | arrivals of the current one
    done = chain[32:]
    grid = int((done != 0).sum())
    assert 1 <= grid <= (n + 511) // 512
    assert word == (T * grid) << 16
    assert (done[:grid] == T * grid).all() and (done[grid:] == 0).all()
    # mixing in other launches (observation variant, general kernel via word0, unchained steps) keeps it consistent
    env.step(acts[0], want_obs=True); ref.step(a_np[0])
    env.step(acts[1], chained=True); ref.step(a_np[1])
    env.step(acts[2], want_status=True, chained=True); ref.step(a_np[2])
    env.step(acts[3], chained=True); ref.step(a_np[3])
    assert env.state_numpy()[0].tobytes() == ref.ships.tobytes()


def test_chained_graph_replay(torch_cuda, default_map):
    """capture_steps chains the captured launches; replays give the oracle's trajectory."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    n, seed, T, max_steps = 20_000, 31, 16, 6
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=max_steps)
    ref = ThreadedOracle(default_map, env.ports, env.port_fuel, env.port_cargo, n, seed, 0, max_steps)
    env.reset()
    acts = torch.randint(0, 4, (T, n), device="cuda", dtype=torch.int16)
    graph = env.capture_steps(acts)
    for rep in range(3):
        graph.replay()
        for t in range(T):
            ref.step(acts[t].cpu().numpy())
        assert env.state_numpy()[0].tobytes() == ref.ships.tobytes(), f"replay {rep}"
    env.step(acts[0], chained=True); ref.step(acts[0].cpu().numpy())      # eager step right after a replay
    assert env.state_numpy()[0].tobytes() == ref.ships.tobytes()


@pytest.mark.parametrize("n,T,policy", [(4096, 64, False), (50_000, 24, False), (1, 40, True), (33, 40, False),
                                        (50_000, 24, True), (1 << 20, 6, False)])
def test_step_many_equals_repeated_step(torch_cuda, default_map, n, T, policy):
    """shipenv_step_many == T x shipenv_step == the oracle: every output of every step."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    seed, max_steps = 0xBEEF, 13
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=max_steps)
    ref = ThreadedOracle(default_map, env.ports, env.port_fuel, env.port_cargo, n, seed, 0, max_steps)
    env.reset()
    env.step_policy(); ref.step(np.full(n, -1, np.int16))          # some cargo on board
    gen = torch.Generator(device="cuda").manual_seed(8)
    acts = torch.randint(0, 4, (T, n), generator=gen, device="cuda", dtype=torch.int16)
    if not policy:                                                 # sprinkle the rest of the flat action space in
        other = torch.randint(-1, 259, (T, n), generator=gen, device="cuda", dtype=torch.int16)
        pick = torch.rand((T, n), generator=gen, device="cuda") < 0.15
        acts = torch.where(pick, other, acts)
    out = env.step_many(None if policy else acts, steps=T, want_status=True, want_taken=True, want_reward64=True)
    a_np = np.full((T, n), -1, np.int16) if policy else acts.cpu().numpy()
    for t in range(T):
        want = ref.step(a_np[t])
        assert (out["reward64"][t].cpu().numpy().view(np.uint64) == want["reward"].view(np.uint64)).all(), t
        assert (out["reward"][t].cpu().numpy() == want["reward"].astype(np.float32)).all()
        assert (out["status"][t].cpu().numpy() == want["status"]).all()
        ok = want["status"] == 0
        assert (out["taken"][t].cpu().numpy()[ok] == want["taken"][ok]).all()
        assert (bits(out["done_bits"][t], n) == want["done"].astype(bool)).all()
        assert (bits(out["reset_bits"][t], n) == want["reset"].astype(bool)).all()
        assert (bits(out["error_bits"][t], n) == (want["status"] != 0)).all()
    g_ships, g_clock = env.state_numpy()
    assert g_ships.tobytes() == ref.ships.tobytes() and (g_clock == ref.clock).all()


@pytest.mark.parametrize("name", ["g1_moves", "g2_policy", "g3_mixed", "g13_free_moves"])
def test_step_many_replays_reference_traces(torch_cuda, name):
    """The reference-generated fixtures (moves, policy, the whole flat action space) through shipenv_step_many in one
    launch: final ships, per-step reward bits and status as recorded from the reference."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    z = golden(name)
    acts = np.ascontiguousarray(z["actions"], dtype=np.int16)
    T, n = acts.shape
    env = BatchedEnvironment(n, passable=z["passable"], ports=z["port_xy"].tolist(), port_fuel=z["port_fuel"].tolist(),
                             port_cargo=z["port_cargo"].tolist(), seed=int(z["seed"]),
                             env_id_base=int(z["env_id_base"]), device="cuda:0", max_steps=int(z["max_steps"]),
                             auto_reset=bool(z["auto_reset"]))
    env.load_state(orc.ships_from(z["init_pos"], z["init_cargo"], z["init_fuel"]), z["init_clock"].astype(np.uint32))
    out = env.step_many(torch.from_numpy(acts).cuda(), want_status=True, want_taken=True, want_reward64=True)
    assert (out["reward64"].cpu().numpy().view(np.uint64) == z["reward"].view(np.uint64)).all()
    assert (out["reward"].cpu().numpy() == z["reward"].astype(np.float32)).all()
    assert (out["status"].cpu().numpy() == z["status"]).all()
    ok = z["status"] == 0
    assert (out["taken"].cpu().numpy()[ok] == z["taken"][ok]).all()
    for t in (0, T // 2, T - 1):
        assert (bits(out["done_bits"][t], n) == z["done"][t].astype(bool)).all()
        assert (bits(out["reset_bits"][t], n) == z["reset"][t].astype(bool)).all()
    ships, clock = env.state_numpy()
    assert ships.tobytes() == orc.ships_from(z["pos"][-1], z["cargo"][-1], z["fuel"][-1]).tobytes()
    assert (clock == z["clock"][-1]).all()
