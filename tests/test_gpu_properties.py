"""GPU: hypothesis-driven edge coverage -- arbitrary ship records (including states no trajectory reaches),
arbitrary int16 actions, clocks near the 16-bit step boundary, all kernel variants -- CUDA vs the oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, example, given, settings, strategies as st

from conftest import golden
from oracle import lib as orc

pytestmark = pytest.mark.gpu

FUEL_EDGES = np.array([0.0, -0.0, 0.5, 0.9, 0.95, 1.0, 1.05, 1.1, 1.1000000000000001, 200.0, -3.5, 1e-300, 1e6,
                       0.8999999999999999, 5e-324, 1e15], np.float64)
ACTION_EDGES = np.array([-1, -1, -1, 0, 1, 2, 3, 4, 8, 9, 10, 29, 58, 59, 60, 75, 258, 259, 260, -2, -32768, 32767,
                         0x4000, 0x4001, 0x4014, 0x5FFF, 0x6000, 0x6001, 0x6010, 0x7FFF, 0x3FFF,
                         # moves by any (dx, dy): -2 - ((dx+90)*181 + (dy+90)); -32762 is the last code, -32763 is junk
                         -2 - (90 * 181 + 90), -2 - (91 * 181 + 91), -2 - (87 * 181 + 92), -2 - (95 * 181 + 83),
                         -2 - (90 * 181 + 89), -2 - (102 * 181 + 90), -32762, -32763, -3], np.int64)


def random_population(rng, n, passable, ports):
    s = np.zeros(n, orc.SHIP_DTYPE)
    on_port = rng.random(n) < 0.35
    k = rng.integers(0, len(ports), n)
    s["x"] = np.where(on_port, np.array(ports)[k, 0], rng.integers(0, 100, n))
    s["y"] = np.where(on_port, np.array(ports)[k, 1], rng.integers(0, 100, n))
    edge = rng.random(n) < 0.15                       # hug the map border so that off-map moves occur
    s["x"] = np.where(edge, rng.choice([0, 99], n), s["x"])
    s["origin"] = rng.choice([0, 1, 2, 3, 4, 255], n)
    s["dest"] = rng.choice([0, 1, 2, 3, 4, 255], n, p=[.18, .18, .18, .18, .18, .1])
    s["cargo"] = rng.choice([0, 0, 1, 2, 7, 25, 49, 50, 51, 60, 1000, 2_000_000, -1, -7], n)
    fuel = rng.random(n) * 210 - 5
    use_edge = rng.random(n) < 0.4
    s["fuel"] = np.where(use_edge, rng.choice(FUEL_EDGES, n), fuel)
    clock = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    near = rng.random(n) < 0.5                        # step counters around max_steps and the 16-bit wrap
    clock = np.where(near, (clock & np.uint32(0xFFFF0000)) | rng.choice([0, 1, 35, 36, 37, 65534, 65535], n).astype(np.uint32), clock)
    return s, clock.astype(np.uint32)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2**63 - 1), n=st.sampled_from([1, 31, 32, 33, 255, 256, 257, 1000, 4099]),
       max_steps=st.sampled_from([0, 1, 37, 65535]), auto_reset=st.booleans(), lean=st.booleans(),
       base=st.sampled_from([0, 7, (1 << 32) - 3, (1 << 40) + 11]), word0=st.sampled_from([0, 0, 1, 5]))
def test_one_step_from_arbitrary_states(seed, n, max_steps, auto_reset, lean, base, word0):
    import torch
    from shippingenv_b200 import BatchedEnvironment, unpack_bits
    z = golden("default_map")
    rng = np.random.default_rng(seed % (1 << 32))
    env = BatchedEnvironment(n, passable=z["passable"], seed=seed, device="cuda:0", env_id_base=base,
                             max_steps=max_steps, auto_reset=auto_reset)
    o = orc.Oracle(z["passable"], env.ports, env.port_fuel, env.port_cargo)
    ships, clock = random_population(rng, n, z["passable"], env.ports)
    acts = np.where(rng.random(n) < 0.5, rng.integers(0, 4, n), rng.choice(ACTION_EDGES, n)).astype(np.int16)
    env.load_state(ships, clock)
    a = torch.from_numpy(acts).cuda()
    if lean:
        r = env.step(a, word0=word0)                                   # fast kernel when word0 == 0
    else:
        r = env.step(a, want_status=True, want_taken=True, want_reward64=True, want_obs=True, word0=word0)
    ref = o.step(ships, clock, acts, seed, base, max_steps, auto_reset, word0)
    g_ships, g_clock = env.state_numpy()
    bad = np.nonzero((g_ships != ships.view(np.uint8).reshape(n, 16)).any(1))[0]
    assert bad.size == 0, (bad[:5], acts[bad[:5]], g_ships[bad[:5]], ships[bad[:5]])
    assert (g_clock == clock).all()
    assert (r.reward.cpu().numpy() == ref["reward"].astype(np.float32)).all()
    assert (unpack_bits(r.done_bits, n).cpu().numpy() == ref["done"].astype(bool)).all()
    assert (unpack_bits(r.reset_bits, n).cpu().numpy() == ref["reset"].astype(bool)).all()
    assert (unpack_bits(r.error_bits, n).cpu().numpy() == (ref["status"] != 0)).all()
    if not lean:
        assert (r.status.cpu().numpy() == ref["status"]).all()
        assert (r.reward64.cpu().numpy().view(np.uint64) == ref["reward"].view(np.uint64)).all()
        ok = ref["status"] == 0
        assert (r.taken.cpu().numpy()[ok] == ref["taken"][ok]).all()
        assert (r.obs.cpu().numpy().view(np.uint32) == o.observe(ships).view(np.uint32)).all()


@settings(max_examples=12, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2**63 - 1), S=st.sampled_from([1, 3, 50]), L=st.sampled_from([0, 1, 17, 100]))
def test_rollouts_from_arbitrary_roots(seed, S, L):
    from shippingenv_b200 import BatchedEnvironment
    z = golden("default_map")
    rng = np.random.default_rng(seed % (1 << 32))
    R = 300
    env = BatchedEnvironment(R, passable=z["passable"], seed=seed, device="cuda:0")
    o = orc.Oracle(z["passable"], env.ports, env.port_fuel, env.port_cargo)
    roots, _ = random_population(rng, R, z["passable"], env.ports)
    env.load_state(roots)
    out = env.rollout(S, L, root_id_base=seed & 0xFFFF, want_final=True)
    want, fin, steps, status = o.rollout(roots, S, L, seed, seed & 0xFFFF)
    assert (out["status"].cpu().numpy() == status).all()
    assert (out["steps"].cpu().numpy() == steps).all()
    assert (out["returns"].cpu().numpy().view(np.uint64) == want.view(np.uint64)).all()
    assert out["final"].cpu().numpy().tobytes() == fin.tobytes()


def random_world(rng, W, H, P, density):
    """Random passability, P distinct port cells (fewer if the map is smaller), random stock."""
    passable = (rng.random((W, H)) < density).astype(np.uint8)
    P = min(P, W * H)
    cells = rng.choice(W * H, size=P, replace=False)
    if P and rng.random() < 0.5:
        cells[0] = rng.choice([0, W * H - 1, H - 1, (W - 1) * H])      # a port in a corner of the map
        cells = np.array(list(dict.fromkeys(cells.tolist())))
    ports = [[int(c) // H, int(c) % H] for c in cells]
    fuel = rng.integers(0, 256, len(ports)).tolist()
    cargo = rng.integers(0, 256, len(ports)).tolist()
    return passable, ports, fuel, cargo


def world_population(rng, n, W, H, ports):
    P = len(ports)
    s = np.zeros(n, orc.SHIP_DTYPE)
    s["x"], s["y"] = rng.integers(0, W, n), rng.integers(0, H, n)
    if P:
        on_port = rng.random(n) < 0.4
        k = rng.integers(0, P, n)
        s["x"] = np.where(on_port, np.array(ports)[k, 0], s["x"])
        s["y"] = np.where(on_port, np.array(ports)[k, 1], s["y"])
    idx = np.arange(P).tolist() + [255]
    s["origin"], s["dest"] = rng.choice(idx, n), rng.choice(idx, n)
    s["cargo"] = rng.choice([0, 0, 1, 7, 49, 50, 51, 300, -2], n)
    s["fuel"] = np.where(rng.random(n) < 0.3, rng.choice(FUEL_EDGES, n), rng.random(n) * 60)
    clock = (rng.integers(0, 1 << 16, n).astype(np.uint32) << np.uint32(16)) | rng.integers(0, 40, n).astype(np.uint32)
    return s, clock


@settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2**63 - 1), W=st.sampled_from([1, 2, 31, 32, 33, 100, 255, 256]),
       H=st.sampled_from([1, 2, 31, 32, 33, 64, 65, 100, 255, 256]), P=st.sampled_from([0, 1, 2, 3, 5, 17, 31, 32]),
       density=st.sampled_from([0.0, 0.3, 0.9, 1.0]), n=st.sampled_from([1, 300, 2049]), lean=st.booleans())
@example(seed=1, W=256, H=256, P=32, density=0.9, n=2049, lean=False)     # largest world, 134-float observation rows
@example(seed=2, W=256, H=256, P=32, density=0.9, n=2049, lean=True)
@example(seed=3, W=33, H=255, P=31, density=0.3, n=300, lean=False)
@example(seed=4, W=1, H=1, P=1, density=0.0, n=300, lean=False)          # the smallest world
@example(seed=5, W=256, H=1, P=2, density=1.0, n=2049, lean=True)        # one column: every N/S move leaves the map
@example(seed=6, W=2, H=256, P=17, density=1.0, n=1, lean=False)
def test_steps_on_arbitrary_worlds(seed, W, H, P, density, n, lean):
    """Map sizes 1..256 on both axes (bitmap rows of 1..8 words), 0..32 ports (observation rows of 6..134 floats),
    ports on the border, all-ground and all-water maps: eight steps of mixed actions, CUDA vs the oracle."""
    import torch
    from shippingenv_b200 import BatchedEnvironment, unpack_bits
    rng = np.random.default_rng(seed % (1 << 32))
    passable, ports, pf, pc = random_world(rng, W, H, P, density)
    P = len(ports)
    auto_reset = P >= 2
    env = BatchedEnvironment(n, passable=passable, ports=ports, port_fuel=pf, port_cargo=pc, seed=seed,
                             device="cuda:0", max_steps=37, auto_reset=auto_reset)
    o = orc.Oracle(passable, ports, pf, pc)
    ships, clock = world_population(rng, n, W, H, ports)
    env.load_state(ships, clock)
    n_act = 4 + P + 250
    for t in range(8):
        kind = rng.random(n)
        acts = np.where(kind < 0.5, rng.integers(0, 4, n), np.where(kind < 0.75, -1, rng.integers(0, n_act + 2, n)))
        # moves by any (dx, dy) (-2 - ((dx+90)*181 + (dy+90))): short ones that mostly stay on the map, and any code
        free = -2 - ((rng.integers(-4, 5, n) + 90) * 181 + (rng.integers(-4, 5, n) + 90))
        acts = np.where(kind > 0.9, free, np.where(kind > 0.88, rng.integers(-32768, -1, n), acts))
        acts = acts.astype(np.int16)
        a = torch.from_numpy(acts).cuda()
        if lean:
            r = env.step(a)
        else:
            r = env.step(a, want_status=True, want_taken=True, want_reward64=True, want_obs=True)
        ref = o.step(ships, clock, acts, seed, 0, 37, auto_reset, 0)
        g_ships, g_clock = env.state_numpy()
        bad = np.nonzero((g_ships != ships.view(np.uint8).reshape(n, 16)).any(1))[0]
        assert bad.size == 0, (t, bad[:5], acts[bad[:5]], g_ships[bad[:5]], ships[bad[:5]])
        assert (g_clock == clock).all(), t
        assert (r.reward.cpu().numpy() == ref["reward"].astype(np.float32)).all(), t
        assert (unpack_bits(r.done_bits, n).cpu().numpy() == ref["done"].astype(bool)).all(), t
        assert (unpack_bits(r.reset_bits, n).cpu().numpy() == ref["reset"].astype(bool)).all(), t
        assert (unpack_bits(r.error_bits, n).cpu().numpy() == (ref["status"] != 0)).all(), t
        if not lean:
            assert (r.status.cpu().numpy() == ref["status"]).all(), t
            assert (r.reward64.cpu().numpy().view(np.uint64) == ref["reward"].view(np.uint64)).all(), t
            assert (r.obs.cpu().numpy().view(np.uint32) == o.observe(ships).view(np.uint32)).all(), t
    if P >= 1:
        assert (env.valid_action_mask().cpu().numpy() == o.valid_actions(ships).astype(bool)).all()


@settings(max_examples=16, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2**63 - 1), W=st.sampled_from([1, 2, 33, 100, 256]), H=st.sampled_from([1, 2, 31, 100, 256]),
       P=st.sampled_from([0, 1, 2, 5, 32]), density=st.sampled_from([0.0, 0.5, 1.0]))
@example(seed=9, W=256, H=256, P=32, density=0.9)
@example(seed=10, W=1, H=2, P=1, density=0.0)
def test_rollouts_and_mcts_on_arbitrary_worlds(seed, W, H, P, density):
    """MCTSAgent._rollout and choose_action from arbitrary roots in arbitrary worlds (including the ones where nothing
    can move and the ones without ports), CUDA vs the oracle."""
    from shippingenv_b200 import BatchedEnvironment
    rng = np.random.default_rng(seed % (1 << 32))
    passable, ports, pf, pc = random_world(rng, W, H, P, density)
    R = 200
    env = BatchedEnvironment(R, passable=passable, ports=ports, port_fuel=pf, port_cargo=pc, seed=seed,
                             device="cuda:0", auto_reset=False)
    o = orc.Oracle(passable, ports, pf, pc)
    roots, _ = world_population(rng, R, W, H, ports)
    env.load_state(roots)
    out = env.rollout(3, 17, root_id_base=5, want_final=True)
    want, fin, steps, status = o.rollout(roots, 3, 17, seed, 5)
    assert (out["status"].cpu().numpy() == status).all()
    assert (out["steps"].cpu().numpy() == steps).all()
    assert (out["returns"].cpu().numpy().view(np.uint64) == want.view(np.uint64)).all()
    assert out["final"].cpu().numpy().tobytes() == fin.tobytes()
    got = {k: v.cpu().numpy() for k, v in env.mcts(12, 10, root_id_base=3).items()}
    ref = orc.mcts(o, roots, 12, 10, seed, 3)
    assert (got["status"] == ref["status"]).all()
    ok = ref["status"] == 0
    for k in ("action", "root_visits", "child_action", "child_visits"):
        assert (got[k][ok] == ref[k][ok]).all(), k
    assert (got["child_total"][ok].view(np.uint64) == ref["child_total"][ok].view(np.uint64)).all()


@settings(max_examples=20, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2**63 - 1), W=st.sampled_from([1, 2, 33, 100, 256]), H=st.sampled_from([1, 31, 100, 256]),
       P=st.sampled_from([0, 1, 2, 5, 32]), word0=st.sampled_from([0, 1, 3, 6]),
       base=st.sampled_from([0, (1 << 32) - 2, (1 << 41) + 5]))
def test_sample_action_and_step_sample_from_arbitrary_states(seed, W, H, P, word0, base):
    """shipenv_sample_action and shipenv_step_sample (the facade's launches) against the oracle on arbitrary states and
    worlds: sampled action, words consumed, status; then the step and the look-ahead draw that follows it."""
    import ctypes
    import torch
    from shippingenv_b200 import BatchedEnvironment, _native as nat, unpack_bits
    rng = np.random.default_rng(seed % (1 << 32))
    passable, ports, pf, pc = random_world(rng, W, H, P, 0.7)
    n = 700
    env = BatchedEnvironment(n, passable=passable, ports=ports, port_fuel=pf, port_cargo=pc, seed=seed, device="cuda:0",
                             env_id_base=base, auto_reset=False)
    o = orc.Oracle(passable, ports, pf, pc)
    ships, clock = world_population(rng, n, W, H, ports)
    env.load_state(ships, clock)
    a, wu, stt = (t.cpu().numpy() for t in env.sample_action(word0=word0))
    ra, rwu, rst = o.sample_action(ships, clock, seed, base, word0)
    assert (stt == rst).all() and (wu == rwu).all() and (a[rst == 0] == ra[rst == 0]).all()
    # step_sample: explicit actions (the sampled ones where there is one, junk elsewhere), then the next draw
    acts = np.where(rst == 0, ra, rng.integers(-3, 270, n)).astype(np.int16)
    d_act = torch.from_numpy(acts).cuda()
    dev = env.device
    nxt = torch.full((n,), -7, dtype=torch.int16, device=dev)
    nwords = torch.zeros(n, dtype=torch.uint8, device=dev)
    nstat = torch.zeros(n, dtype=torch.uint8, device=dev)
    reward64 = torch.zeros(n, dtype=torch.float64, device=dev)
    status = torch.zeros(n, dtype=torch.uint8, device=dev)
    io = nat.StepIO()
    io.reward_f64, io.status = reward64.data_ptr(), status.data_ptr()
    io.done_bits, io.reset_bits, io.error_bits = env.done_bits.data_ptr(), env.reset_bits.data_ptr(), env.error_bits.data_ptr()
    env.ctx.step_sample(env.ships.data_ptr(), env.clock.data_ptr(), d_act.data_ptr(), io, nxt.data_ptr(), nwords.data_ptr(),
                        nstat.data_ptr(), seed, base, n, 0, 0, int(wu.max()) if n else 0, torch.cuda.current_stream().cuda_stream)
    w0 = int(wu.max())
    ref = o.step(ships, clock, acts, seed, base, 0, False, w0)
    g_ships, g_clock = env.state_numpy()
    assert (g_ships == ships.view(np.uint8).reshape(n, 16)).all() and (g_clock == clock).all()
    assert (status.cpu().numpy() == ref["status"]).all()
    assert (reward64.cpu().numpy().view(np.uint64) == ref["reward"].view(np.uint64)).all()
    assert (unpack_bits(env.done_bits, n).cpu().numpy() == ref["done"].astype(bool)).all()
    assert (unpack_bits(env.error_bits, n).cpu().numpy() == (ref["status"] != 0)).all()
    # the look-ahead: word 0 of the new cell after an accepted step, word w0 of the same cell after a rejected one
    ok = ref["status"] == 0
    na0, nw0, ns0 = o.sample_action(ships, clock, seed, base, 0)
    na1, nw1, ns1 = o.sample_action(ships, clock, seed, base, w0)
    want_a, want_w, want_s = np.where(ok, na0, na1), np.where(ok, nw0, nw1), np.where(ok, ns0, ns1)
    assert (nstat.cpu().numpy() == want_s).all() and (nwords.cpu().numpy() == want_w).all()
    assert (nxt.cpu().numpy()[want_s == 0] == want_a[want_s == 0]).all()
