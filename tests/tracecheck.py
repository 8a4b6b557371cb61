"""Replay a golden trace (tests/golden/*.npz, produced by the unmodified reference) through a
stepper and compare every field bit-for-bit.  Two steppers exist: the C oracle (CPU tests) and
the CUDA path through the C ABI (GPU tests)."""
import numpy as np

from oracle import lib as orc

FIELDS = ("pos", "cargo", "fuel", "reward", "done", "status", "reset", "taken", "clock")


class OracleStepper:
    def __init__(self, z):
        self.o = orc.Oracle.from_npz(z)
        self.seed, self.base = int(z["seed"]), int(z["env_id_base"])

    def load(self, ships, clock):
        self.ships, self.clock = ships.copy(), clock.copy()

    def step(self, actions, max_steps, auto_reset, word0):
        out = self.o.step(self.ships, self.clock, actions, self.seed, self.base, max_steps, auto_reset, word0)
        return out, self.ships.copy(), self.clock.copy()


class CudaStepper:
    def __init__(self, z):
        import torch
        from shippingenv_b200 import BatchedEnvironment
        self.torch = torch
        n = z["init_cargo"].shape[0]
        self.env = BatchedEnvironment(
            n, passable=z["passable"], ports=z["port_xy"].tolist(), port_fuel=z["port_fuel"].tolist(),
            port_cargo=z["port_cargo"].tolist(), seed=int(z["seed"]), env_id_base=int(z["env_id_base"]),
            device="cuda:0")

    def load(self, ships, clock):
        self.env.load_state(ships, clock)

    def step(self, actions, max_steps, auto_reset, word0):
        env, torch = self.env, self.torch
        env.max_steps, env.auto_reset = max_steps, auto_reset
        a = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.int16)).to(env.device)
        r = env.step(a, want_status=True, want_taken=True, want_reward64=True, word0=word0)
        from shippingenv_b200 import unpack_bits
        out = dict(
            reward=r.reward64.cpu().numpy(), reward32=r.reward.cpu().numpy(),
            done=unpack_bits(r.done_bits, env.n).cpu().numpy().astype(np.uint8),
            reset=unpack_bits(r.reset_bits, env.n).cpu().numpy().astype(np.uint8),
            error=unpack_bits(r.error_bits, env.n).cpu().numpy().astype(np.uint8),
            status=r.status.cpu().numpy(), taken=r.taken.cpu().numpy())
        ships, clock = env.state_numpy()
        return out, ships.view(orc.SHIP_DTYPE).reshape(-1), clock


def replay(z, stepper):
    """Assert the stepper reproduces trace z exactly.  Returns the number of env-steps checked."""
    ships = orc.ships_from(z["init_pos"], z["init_cargo"], z["init_fuel"])
    stepper.load(ships, z["init_clock"].astype(np.uint32))
    T = z["actions"].shape[0]
    max_steps, auto_reset, word0 = int(z["max_steps"]), bool(z["auto_reset"]), int(z["word0"])
    for t in range(T):
        out, s, clock = stepper.step(z["actions"][t], max_steps, auto_reset, word0)
        got = dict(
            pos=np.stack([s["x"], s["y"], s["origin"], s["dest"]], axis=1), cargo=s["cargo"],
            fuel=s["fuel"].view(np.uint64), reward=out["reward"].view(np.uint64), done=out["done"],
            status=out["status"], reset=out["reset"], taken=out["taken"], clock=clock)
        for k in FIELDS:
            want = z[k][t]
            if k == "reward":
                want = want.view(np.uint64)
            if k == "taken":
                ok = out["status"] == 0           # `taken` is only defined for accepted steps
                bad = np.nonzero((got[k] != want) & ok)[0]
            else:
                bad = np.nonzero((got[k] != want).reshape(len(want), -1).any(axis=1))[0]
            assert bad.size == 0, (f"step {t} field {k}: env {bad[0]} got {got[k][bad[0]]} want {want[bad[0]]}"
                                   f" (action {z['actions'][t][bad[0]]})")
        if "reward32" in out:   # the f32 output is the fp64 reward rounded once
            assert (out["reward32"] == z["reward"][t].astype(np.float32)).all(), f"step {t}: f32 reward"
            assert (out["error"] == (z["status"][t] != 0)).all(), f"step {t}: error ballot"
    return T * len(ships)


TRACES = ["g1_moves", "g2_policy", "g3_mixed", "g4_navigator", "g5_edges_moves", "g5_edges_policy",
          "g6_crafted", "g7_word0", "g13_free_moves"]
