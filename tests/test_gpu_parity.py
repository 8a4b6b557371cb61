"""GPU: the CUDA path, called through the C ABI, against the reference-generated fixtures and
against the CPU oracle on seeded inputs.  Bit-exact for positions, ports, cargo, fuel (f64 bits),
done/reset/status, clock and the f64 reward; the f32 reward is the f64 one rounded once."""
import numpy as np
import pytest

from conftest import golden
from oracle import lib as orc
from tracecheck import TRACES, CudaStepper, replay

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("name", TRACES)
def test_cuda_replays_reference_trace(torch_cuda, name):
    z = golden(name)
    assert replay(z, CudaStepper(z)) > 0


def test_cuda_reset_matches_reference(torch_cuda):
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g1_moves")
    n = z["init_cargo"].shape[0]
    env = BatchedEnvironment(n, passable=z["passable"], seed=int(z["seed"]), device="cuda:0")
    env.reset()
    ships, clock = env.state_numpy()
    want = orc.ships_from(z["init_pos"], z["init_cargo"], z["init_fuel"])
    assert ships.tobytes() == want.tobytes()
    assert (clock == 65536).all()


def test_cuda_masked_reset(torch_cuda):
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    z = golden("default_map")
    n = 100
    env = BatchedEnvironment(n, passable=z["passable"], device="cuda:0")
    env.reset()
    before, clk0 = env.state_numpy()
    env.step_policy()
    mid, clk1 = env.state_numpy()
    mask = np.zeros(4, np.uint32)
    chosen = [0, 5, 31, 32, 64, 99]
    for i in chosen:
        mask[i >> 5] |= np.uint32(1 << (i & 31))
    env.reset(torch.from_numpy(mask.view(np.int32)).to(env.device))
    after, clk2 = env.state_numpy()
    o = orc.Oracle(z["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = mid.view(orc.SHIP_DTYPE).reshape(-1).copy(), clk1.copy()
    m8 = np.zeros(n, np.uint8); m8[chosen] = 1
    o.reset(ships, clock, env.seed, 0, m8)
    assert after.tobytes() == ships.tobytes() and (clk2 == clock).all()
    assert (clk2[chosen] == 2 << 16).all() and (np.delete(clk2, chosen) == (1 << 16) + 1).all()


@pytest.mark.parametrize("n,policy_every", [(1, 2), (33, 3), (4096, 4), (50_000, 0)])
def test_cuda_vs_oracle_seeded(torch_cuda, default_map, n, policy_every):
    """Config 2 of BASELINE.json (4,096 envs, default map and ports, random actions, auto-reset),
    and ragged sizes around it, against the oracle for 256 steps."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment, unpack_bits
    seed, max_steps = 0x5EED, 1000 if n == 4096 else 150
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0",
                             max_steps=max_steps, auto_reset=True)
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    env.reset(); o.reset(ships, clock, seed)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    T = 256 if n <= 4096 else 64
    n_done = 0
    for t in range(T):
        a = torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.int16)
        if policy_every and t % policy_every == 0:
            a.fill_(-1)
        r = env.step(a, want_status=True, want_reward64=True, want_taken=True)
        ref = o.step(ships, clock, a.cpu().numpy(), seed, 0, max_steps, True)
        g_ships, g_clock = env.state_numpy()
        assert g_ships.tobytes() == ships.tobytes(), f"ships differ at step {t}"
        assert (g_clock == clock).all()
        assert (r.reward64.cpu().numpy().view(np.uint64) == ref["reward"].view(np.uint64)).all()
        assert (r.reward.cpu().numpy() == ref["reward"].astype(np.float32)).all()
        assert (r.status.cpu().numpy() == ref["status"]).all()
        assert (unpack_bits(r.done_bits, n).cpu().numpy() == ref["done"].astype(bool)).all()
        assert (unpack_bits(r.reset_bits, n).cpu().numpy() == ref["reset"].astype(bool)).all()
        ok = ref["status"] == 0
        assert (r.taken.cpu().numpy()[ok] == ref["taken"][ok]).all()
        n_done += int(ref["done"].sum())
    if n == 4096:
        assert n_done > 0      # some ships run out of fuel within 256 steps


def test_cuda_observation(torch_cuda):
    """Fused observation epilogue and the stand-alone observe kernel == preprocess_state rows."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g2_policy")
    n = 64
    env = BatchedEnvironment(n, passable=z["passable"], seed=int(z["seed"]), device="cuda:0")
    env.load_state(orc.ships_from(z["pos"][-1], z["cargo"][-1], z["fuel"][-1]), z["clock"][-1])
    obs = env.observe().cpu().numpy()
    assert (obs.view(np.uint32) == z["obs"].view(np.uint32)).all()
    # fused: step with observation output for ragged sizes, compared with the oracle's rows
    for n in (1, 7, 255, 256, 257, 1000, 5001):
        env = BatchedEnvironment(n, passable=z["passable"], seed=7, device="cuda:0", max_steps=40)
        o = orc.Oracle(z["passable"], env.ports, env.port_fuel, env.port_cargo)
        ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
        env.reset(); o.reset(ships, clock, 7)
        for t in range(50):
            r = env.step_policy(want_obs=True)
            o.step(ships, clock, np.full(n, -1, np.int16), 7, 0, 40, True)
            want = o.observe(ships)
            got = r.obs.cpu().numpy()
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), f"n={n} step {t}"
        assert env.state_numpy()[0].tobytes() == ships.tobytes()
        assert (env.observe().cpu().numpy() == want).all()


@pytest.mark.parametrize("L", [100, 25])
def test_cuda_rollout_matches_reference_rollout(torch_cuda, L):
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g8_rollouts")
    roots = orc.ships_from(z["root_pos"], z["root_cargo"], z["root_fuel"])
    R, S = z[f"ret_L{L}"].shape
    env = BatchedEnvironment(R, passable=z["passable"], seed=int(z["seed"]), device="cuda:0")
    env.load_state(roots)
    out = env.rollout(S, L, root_id_base=int(z["root_id_base"]), want_final=True)
    assert (out["status"].cpu().numpy() == 0).all()
    assert (out["returns"].cpu().numpy().view(np.uint64) == z[f"ret_L{L}"]).all()
    fin = out["final"].cpu().numpy().view(orc.SHIP_DTYPE).reshape(R, S)
    assert (np.stack([fin["x"], fin["y"], fin["origin"], fin["dest"]], -1) == z[f"fin_pos_L{L}"]).all()
    assert (fin["cargo"] == z[f"fin_cargo_L{L}"]).all()
    assert (fin["fuel"].view(np.uint64) == z[f"fin_fuel_L{L}"]).all()


def test_cuda_rollout_vs_oracle_large(torch_cuda, default_map):
    """Config 4 shape at a size the oracle finishes in seconds: 2,048 roots x 50 sims x 100 steps."""
    from shippingenv_b200 import BatchedEnvironment
    R, S, L, seed = 2048, 50, 100, 99
    env = BatchedEnvironment(R, passable=default_map["passable"], seed=seed, device="cuda:0", auto_reset=False)
    env.reset()
    for _ in range(10):
        env.step_policy()
    roots, _ = env.state_numpy()
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    want, fin, steps, status = o.rollout(roots.view(orc.SHIP_DTYPE).reshape(-1), S, L, seed)
    out = env.rollout(S, L, want_final=True)
    assert (out["returns"].cpu().numpy().view(np.uint64) == want.view(np.uint64)).all()
    assert out["final"].cpu().numpy().tobytes() == fin.tobytes()
    assert (out["steps"].cpu().numpy() == steps).all()


def test_cuda_sample_action(torch_cuda, default_map):
    from shippingenv_b200 import BatchedEnvironment
    n, seed = 3000, 5
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", auto_reset=False)
    env.reset()
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    for t in range(20):
        a, words, st = env.sample_action()
        # sample_action() then step(a) with word0 = words consumed == one policy step when accepted
        ships, clock = env.state_numpy()
        ships = ships.view(orc.SHIP_DTYPE).reshape(-1).copy(); clock = clock.copy()
        ref = o.step(ships, clock, np.full(n, -1, np.int16), seed, 0, 0, False)
        assert (st.cpu().numpy() == 0).all() and (words.cpu().numpy() == 1).all()
        r = env.step(a, want_status=True, word0=1)
        acc = r.status.cpu().numpy() == 0
        # where the first draw was accepted, the policy step took the same action
        assert (a.cpu().numpy()[acc] == ref["taken"][acc]).mean() > 0.99
        same = acc & (a.cpu().numpy() == ref["taken"])
        got, _ = env.state_numpy()
        assert (got.view(orc.SHIP_DTYPE).reshape(-1)[same] == ships[same]).all()


def test_cuda_consts_can_vary(torch_cuda, default_map):
    """The constants struct reaches the kernel: different rewards / fuel / capacity, still oracle-exact."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    consts = dict(fuel0=37.5, cargo0=3, capacity=20, r_deliver=1.25, r_arrive=7.0, r_closer=0.3,
                  p_no_fuel=-4.0, p_ground=-0.7, p_water=-0.11, p_loss=-1.3, p_farther=-0.9, p_use_fuel=-0.003,
                  r_take_fuel=0.4, r_take_cargo=0.6)
    n, seed = 2000, 11
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=90,
                             consts=consts)
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo, **consts)
    ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    env.reset(); o.reset(ships, clock, seed)
    for t in range(200):
        r = env.step_policy(want_reward64=True)
        ref = o.step(ships, clock, np.full(n, -1, np.int16), seed, 0, 90, True)
        assert env.state_numpy()[0].tobytes() == ships.tobytes(), t
        assert (r.reward64.cpu().numpy().view(np.uint64) == ref["reward"].view(np.uint64)).all(), t


def test_cuda_episode_statistics(torch_cuda, default_map):
    """On-device episode return / length accumulators against the same quantities computed from the oracle."""
    from shippingenv_b200 import BatchedEnvironment
    n, seed, max_steps, T = 3000, 21, 37, 150
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=max_steps,
                             track_stats=True)
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    env.reset(); o.reset(ships, clock, seed)
    ep = np.zeros(n)
    rets, lens = [], []
    for t in range(T):
        before = clock.copy()
        env.step_policy()
        ref = o.step(ships, clock, np.full(n, -1, np.int16), seed, 0, max_steps, True)
        ok = ref["status"] == 0
        ep[ok] += ref["reward"][ok]
        ended = ref["reset"] == 1
        rets += ep[ended].tolist()
        lens += ((before[ended] + 1) & 0xFFFF).tolist()
        ep[ended] = 0.0
    s = env.episode_summary()
    assert s["total_episodes"] == len(rets) > n and s["total_steps"] == sum(lens)
    assert abs(s["mean_reward"] - np.mean(rets)) < 1e-9 and abs(s["std_reward"] - np.std(rets)) < 1e-6
    assert s["max_reward"] == max(rets) and s["min_reward"] == min(rets)
    assert np.allclose(env.ep_return.cpu().numpy(), ep, rtol=0, atol=1e-12)
    assert env.state_numpy()[0].tobytes() == ships.tobytes()


def test_cuda_full_size_properties(torch_cuda, default_map):
    """BASELINE config 3 size (2,097,152 envs per GPU): properties that do not need the oracle.
    (a) sharding invariance: one launch over all envs == four launches over contiguous quarters with the
        matching env_id_base (what ranks do), bit for bit, through fast and general kernels alike;
    (b) invariants of the transition: ships stay on passable cells inside the map, the clock advances by one
        per accepted step or jumps to the next episode, rejected envs are untouched, a reset puts the ship on
        its origin port with full fuel, fuel never increases under MOVE actions without a reset."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment, unpack_bits
    n, seed, T = 1 << 21, 0x5EED, 24
    passable = default_map["passable"]
    whole = BatchedEnvironment(n, passable=passable, seed=seed, device="cuda:0", max_steps=17)
    q = n // 4
    parts = [BatchedEnvironment(q, passable=passable, seed=seed, device="cuda:0", max_steps=17, env_id_base=k * q)
             for k in range(4)]
    whole.reset()
    for p in parts:
        p.reset()
    gen = torch.Generator(device="cuda").manual_seed(7)
    pas = torch.from_numpy(passable.astype(np.bool_)).cuda()
    ports = torch.tensor(whole.ports, device="cuda")
    for t in range(T):
        a = torch.randint(0, 4, (n,), generator=gen, device="cuda", dtype=torch.int16)
        if t % 5 == 4:
            a = torch.randint(-1, 259, (n,), generator=gen, device="cuda", dtype=torch.int16)
        before_ships, before_clock = whole.ships.clone(), whole.clock.clone()
        r = whole.step(a, want_status=(t % 2 == 0))          # alternate general / fast kernel
        for k, p in enumerate(parts):
            p.step(a[k * q:(k + 1) * q].contiguous(), want_status=(t % 2 == 1))
        assert torch.equal(whole.ships, torch.cat([p.ships for p in parts])), f"split differs at step {t}"
        assert torch.equal(whole.clock, torch.cat([p.clock for p in parts]))
        assert torch.equal(whole.reward, torch.cat([p.reward for p in parts]))
        assert torch.equal(whole.done_bits, torch.cat([p.done_bits for p in parts]))
        assert torch.equal(whole.reset_bits, torch.cat([p.reset_bits for p in parts]))
        err = unpack_bits(r.error_bits, n)
        rst = unpack_bits(r.reset_bits, n)
        x, y = whole.x.long(), whole.y.long()
        assert bool((x < 100).all()) and bool((y < 100).all()) and bool(pas[x, y].all())
        assert torch.equal(whole.ships[err], before_ships[err]) and torch.equal(whole.clock[err], before_clock[err])
        ok = ~err & ~rst
        assert bool((whole.clock[ok] == before_clock[ok] + 1).all())
        assert bool((whole.clock[rst] & 0xFFFF == 0).all())
        assert bool(((whole.clock[rst] >> 16) == (before_clock[rst] >> 16) + 1).all())
        o = whole.origin.long()[rst]
        assert bool((whole.fuel[rst] == 200.0).all()) and bool((whole.cargo[rst] == 0).all())
        assert torch.equal(torch.stack([x[rst], y[rst]], 1), ports[o])
        if t % 5 != 4:
            fuel_before = before_ships.view(torch.float64)[:, 1]
            assert bool((whole.fuel[ok] <= fuel_before[ok]).all())
        assert bool((whole.reward[err] == 0).all())


def test_cuda_valid_action_mask(torch_cuda, default_map):
    """The DQN valid-action mask on device == DQNAgent.is_valid_action of the reference (fixture g9) and ==
    the oracle on a larger seeded population."""
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g9_valid_actions")
    n = len(z["cargo"])
    env = BatchedEnvironment(n, passable=z["passable"], device="cuda:0")
    env.load_state(orc.ships_from(z["pos"], z["cargo"], z["fuel"]))
    mask = env.valid_action_mask().cpu().numpy()
    assert mask.shape == (n, 259) and (mask == z["valid"].astype(bool)).all()
    n = 20_000
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=3, device="cuda:0", max_steps=30)
    env.reset()
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    for t in range(12):
        env.step_policy()
        ships, _ = env.state_numpy()
        want = o.valid_actions(ships.view(orc.SHIP_DTYPE).reshape(-1))
        assert (env.valid_action_mask().cpu().numpy() == want.astype(bool)).all(), t


def test_cuda_discretize(torch_cuda):
    """discretize_state keys on device == the reference's (fixture g10) == the oracle on random records."""
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g10_discretize")
    n = len(z["cargo"])
    env = BatchedEnvironment(n, passable=golden("default_map")["passable"], device="cuda:0")
    env.load_state(orc.ships_from(z["pos"], z["cargo"], z["fuel"]))
    assert (env.discretize().cpu().numpy() == z["keys"]).all()
    rng = np.random.default_rng(1)
    n = 100_000
    s = np.zeros(n, orc.SHIP_DTYPE)
    s["x"], s["y"] = rng.integers(0, 100, n), rng.integers(0, 100, n)
    s["origin"], s["dest"] = rng.choice([0, 1, 2, 3, 4, 255], n), rng.choice([0, 1, 2, 3, 4, 255], n)
    s["fuel"] = np.where(rng.random(n) < 0.5, rng.random(n) * 260 - 60, rng.integers(-30, 30, n) * 10.0)
    env = BatchedEnvironment(n, passable=golden("default_map")["passable"], device="cuda:0")
    env.load_state(s)
    assert (env.discretize().cpu().numpy() == orc.discretize(s)).all()


@pytest.mark.parametrize("S,L", [(50, 100), (12, 20)])
def test_cuda_mcts_matches_reference_choose_action(torch_cuda, S, L):
    """The on-device MCTS (tree + rollouts in one kernel) against the unmodified reference agent (fixture g11)."""
    from shippingenv_b200 import BatchedEnvironment
    z = golden("g11_mcts")
    roots = orc.ships_from(z["root_pos"], z["root_cargo"], z["root_fuel"])
    env = BatchedEnvironment(len(roots), passable=z["passable"], seed=int(z["seed"]), device="cuda:0")
    env.load_state(roots)
    out = {k: v.cpu().numpy() for k, v in env.mcts(S, L, root_id_base=int(z["root_id_base"])).items()}
    tag = f"S{S}_L{L}"
    assert (out["status"] == 0).all()
    assert (out["action"] == z[f"action_{tag}"]).all()
    assert (out["root_visits"] == z[f"root_visits_{tag}"]).all()
    assert (out["child_action"] == z[f"child_action_{tag}"]).all()
    assert (out["child_visits"] == z[f"child_visits_{tag}"]).all()
    assert (out["child_total"].view(np.uint64) == z[f"child_total_{tag}"]).all()


def test_cuda_mcts_vs_oracle_large(torch_cuda, default_map):
    """4,096 roots x 50 simulations x 100-step rollouts (BASELINE config 4 shape) against the oracle, plus roots
    that the tree cannot handle (no destination) and a root without any legal expansion."""
    from shippingenv_b200 import BatchedEnvironment
    R, seed = 4096, 4242
    env = BatchedEnvironment(R, passable=default_map["passable"], seed=seed, device="cuda:0", auto_reset=False)
    env.reset()
    for _ in range(12):
        env.step_policy()
    roots, _ = env.state_numpy()
    roots = roots.view(orc.SHIP_DTYPE).reshape(-1).copy()
    roots["dest"][:3] = 255                         # unsupported node state -> status 10
    env.load_state(roots)
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    want = orc.mcts(o, roots, 50, 100, seed)
    got = {k: v.cpu().numpy() for k, v in env.mcts(50, 100).items()}
    assert (got["status"] == want["status"]).all() and (got["status"][:3] == 10).all() and (got["status"][3:] == 0).all()
    for k in ("action", "root_visits", "child_action", "child_visits"):
        assert (got[k] == want[k]).all(), k
    assert (got["child_total"].view(np.uint64) == want["child_total"].view(np.uint64)).all()
    assert len(np.unique(got["action"][3:])) == 4


@pytest.mark.parametrize("n,max_steps", [(1, 50), (300, 50), (70_001, 120)])
def test_cuda_policy_fast_kernel(torch_cuda, default_map, n, max_steps):
    """SHIPENV_F_ALL_POLICY through the throughput kernel (reward_f32 + masks only): sample_action + step in
    the straight-line path, resets / cargo pick-up / loss / arrival / off-map re-draws through the deferred
    general path -- every step bit-identical to the oracle's policy step."""
    from shippingenv_b200 import BatchedEnvironment, unpack_bits
    seed = 909
    env = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=max_steps)
    o = orc.Oracle(default_map["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    env.reset(); o.reset(ships, clock, seed)
    pol = np.full(n, -1, np.int16)
    n_reset = n_arrive = 0
    for t in range(260 if n < 1000 else 140):
        r = env.step_policy()
        ref = o.step(ships, clock, pol, seed, 0, max_steps, True)
        g_ships, g_clock = env.state_numpy()
        assert g_ships.tobytes() == ships.tobytes(), f"ships differ at step {t}"
        assert (g_clock == clock).all()
        assert (r.reward.cpu().numpy() == ref["reward"].astype(np.float32)).all()
        assert (unpack_bits(r.done_bits, n).cpu().numpy() == ref["done"].astype(bool)).all()
        assert (unpack_bits(r.reset_bits, n).cpu().numpy() == ref["reset"].astype(bool)).all()
        assert (unpack_bits(r.error_bits, n).cpu().numpy() == (ref["status"] != 0)).all()
        n_reset += int(ref["reset"].sum()); n_arrive += int((ref["reward"] > 5).sum())
    assert n_reset > 0
    if n > 1000:
        assert n_arrive > 0


def test_cuda_graph_replay_of_steps(torch_cuda, default_map):
    """BASELINE configs[1] (4,096 envs): T steps captured in a CUDA graph == the same T steps launched one by one,
    and a second replay keeps going exactly like T more eager steps (no host-side state in the launches)."""
    torch = torch_cuda
    from shippingenv_b200 import BatchedEnvironment
    n, T, seed = 4096, 24, 31
    gen = torch.Generator(device="cuda").manual_seed(5)
    acts = torch.randint(0, 4, (T, n), generator=gen, device="cuda", dtype=torch.int16)
    eager = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=30)
    graphed = BatchedEnvironment(n, passable=default_map["passable"], seed=seed, device="cuda:0", max_steps=30)
    eager.reset(); graphed.reset()
    graphed.step(acts[0]); eager.step(acts[0])                # warm the launch path outside the capture
    state0 = graphed.ships.clone(), graphed.clock.clone()
    g = graphed.capture_steps(acts)
    graphed.ships.copy_(state0[0]); graphed.clock.copy_(state0[1])   # the capture itself does not run the steps
    for rep in range(2):
        g.replay()
        for t in range(T):
            eager.step(acts[t])
        torch.cuda.synchronize()
        assert torch.equal(eager.ships, graphed.ships) and torch.equal(eager.clock, graphed.clock), rep
        assert torch.equal(eager.reward, graphed.reward) and torch.equal(eager.reset_bits, graphed.reset_bits)


@pytest.mark.parametrize("n", [1, 1000, 600_001, 2_097_152 + 700, 4_500_000])
def test_cuda_step_host_equals_step(torch_cuda, default_map, n):
    """shipenv_step_host (host actions in, host reward + masks out, chunks pipelined on internal streams) gives what
    shipenv_step gives on the same states; page-locked and pageable host buffers both work."""
    import torch
    from shippingenv_b200 import BatchedEnvironment
    envs = [BatchedEnvironment(n, passable=default_map["passable"], seed=77, device="cuda:0", max_steps=40,
                               env_id_base=5) for _ in range(2)]
    for e in envs:
        e.reset()
    gen = torch.Generator().manual_seed(n)
    for t in range(6):
        acts = torch.randint(0, 4, (n,), generator=gen, dtype=torch.int16)
        if t % 3 == 2:
            acts[::7] = -1                                     # policy actions travel through the host path as well
            acts[3::11] = 4 + (t % 5)
        want = envs[0].step(acts.cuda())
        host_acts = acts.pin_memory() if t % 2 == 0 else acts
        if t == 4:      # caller-provided pageable outputs
            outs = (torch.empty(n, dtype=torch.float32), *(torch.empty((n + 31) // 32, dtype=torch.int32) for _ in range(3)))
            got = envs[1].step_host(host_acts, *outs, chunks=3)
        else:
            got = envs[1].step_host(host_acts, chunks=t % 2)
        torch.cuda.synchronize()
        assert torch.equal(got.reward, want.reward.cpu()), t
        assert torch.equal(got.done_bits, want.done_bits.cpu()), t
        assert torch.equal(got.reset_bits, want.reset_bits.cpu()), t
        assert torch.equal(got.error_bits, want.error_bits.cpu()), t
        assert torch.equal(envs[0].ships, envs[1].ships) and torch.equal(envs[0].clock, envs[1].clock), t


def test_soak_mixed_actions_and_launch_kinds(torch_cuda):
    """tests/soak_check.py, short: 200,001 envs x 96 steps of a mixed action stream (the four moves, the random policy,
    the whole flat action space, moves of any (dx, dy), junk codes) with auto-reset at 9 steps, through plain, chained,
    lean and many-step launches in random alternation; every reward (f64 bits), status, taken action and mask of every
    step and the state every few steps against the C oracle.  (Run at 1,048,653 envs x 192 steps and 300,001 x 258 on a
    B200 at the end of round 2: 2.8e8 env-steps identical.)"""
    from soak_check import soak
    assert soak(envs=200_001, steps=96, seed=3, max_steps=9) >= 200_001 * 96
