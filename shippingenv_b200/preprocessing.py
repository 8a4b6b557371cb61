"""Host-side agent utilities with the reference's names (utils/preprocessing.py), for the N=1
facade.  The batched path does the same work on the device: ``shipenv_observe`` / the fused
observation epilogue write ``preprocess_state`` rows, and the kernels decode flat action indices
exactly like ``map_action_to_env_action``."""
import numpy as np

from . import _native as nat
from .types import ActionType, MOVES

N_MOVES = 4


def preprocess_state(env_state):
    """State dict -> array [1, 6+4P]: x, y, fuel, "cargo", origin|-1, dest|-1, then (px, py, fuel, cargo) per port
    (utils/preprocessing.py:25-62)."""
    ship = env_state["ship"]
    none_to = lambda v: -1 if v is None else v
    row = [ship["position"][0], ship["position"][1], ship["fuel"], ship["cargo"],
           none_to(ship["origin_port_index"]), none_to(ship["destination_port_index"])]
    for port in env_state["ports"]:
        row += [port["position"][0], port["position"][1], port["fuel"], port["cargo"]]
    return np.reshape(row, [1, len(row)])


def discretize_state(state):
    """Hashable key of the tabular agents (utils/preprocessing.py:65-90): position, two coarse buckets, ports."""
    ship = state["ship"]
    none_to = lambda v: -1 if v is None else v
    return (tuple(ship["position"]), min(int(ship["cargo"] / 10), 4), min(int(ship["fuel"] / 20), 4),
            none_to(ship["origin_port_index"]), none_to(ship["destination_port_index"]))


def get_action_space_size(env):
    """4 moves + P ports + 50 cargo amounts + 200 fuel amounts (utils/preprocessing.py:93-108)."""
    return N_MOVES + len(env.port_positions) + nat.N_CARGO_ACTIONS + nat.N_FUEL_ACTIONS


def map_action_to_env_action(agent_action, env):
    """Flat index -> [category, value] (utils/preprocessing.py:111-137); the kernels use the same split."""
    P = len(env.port_positions)
    if agent_action < N_MOVES:
        return [ActionType.MOVE_SHIP, MOVES[agent_action]]
    if agent_action < N_MOVES + P:
        return [ActionType.SELECT_PORT, agent_action - N_MOVES]
    if agent_action < N_MOVES + P + nat.N_CARGO_ACTIONS:
        return [ActionType.TAKE_CARGO, agent_action - (N_MOVES + P)]
    return [ActionType.TAKE_FUEL, agent_action - (N_MOVES + P + nat.N_CARGO_ACTIONS)]


def move_action(dx, dy):
    """Flat int16 action of MOVE_SHIP by (dx, dy) for the batched API: 0..3 for the four moves of ShipMove, the
    SHIPENV_ACT_MOVE_XY code (include/shipenv.h) for any other pair with |dx|, |dy| <= 90 -- the reference's
    _move_ship takes any pair (shipping/environment.py:278-287)."""
    key = (int(dx), int(dy))
    if key in MOVES:
        return MOVES.index(key)
    if max(abs(key[0]), abs(key[1])) > nat.MOVE_XY_MAX:
        raise ValueError(f"moves of more than {nat.MOVE_XY_MAX} cells per axis have no action code")
    return nat.ACT_MOVE_XY_BASE - ((key[0] + 90) * 181 + (key[1] + 90))


def find_current_port_index(state):
    pos = tuple(state["ship"]["position"])
    for idx, port in enumerate(state["ports"]):
        if tuple(port["position"]) == pos:
            return idx
    return None


def get_state_size(env):
    return (1, preprocess_state(env.reset()).shape[1])
