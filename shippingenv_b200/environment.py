"""``Environment``: the reference's single-ship Python API (shipping/environment.py:28-376) on the
sm_100a kernels.  It is the N=1 view of the batched path: every transition, reset and
``sample_action`` is computed by ``libshipenv.so`` on the GPU; this class only keeps the Python
attributes the agents read and write and turns status codes back into the reference's exceptions.

Contract: the attributes are the source of truth.  Agents assign them wholesale
(agents/mcts.py:199-208), so before every launch the ship record is packed from the attributes
into a page-locked host block, the kernel reads and writes that block in place (zero-copy over
PCIe: one launch + one stream synchronisation per call), and the attributes are refreshed from it.
The world (map + port table) is re-uploaded only when it changed.

Differences from the reference, all documented in DESIGN.md: draws come from the Philox word
stream (seeded from ``random`` at construction unless ``seed`` is given) instead of MT19937;
moves may be any integer pair with ``|dx|, |dy| <= 90`` (the int16 action code has no room for
more); take amounts must be integral; ``reset`` with fewer than two ports raises instead of
looping forever; ``sample_action`` with ``fuel == 0`` at a port raises ``TypeError`` like the
reference does, but never consumes a draw.
"""
import os
import random as _pyrandom
import struct

import numpy as np
import torch

from . import _native as nat
from .mapio import derive_grid
from .types import ActionType, Color, Entity, Initial, MOVE_INDEX, MOVES, Penalty, Reward, ShipMove  # noqa: F401

_ERRORS = {
    nat.ST_OFF_MAP: (ValueError, "Move is out of range"),                                     # :284
    nat.ST_NO_DEST: (Exception, "Cannot move without destination port"),                      # :276
    nat.ST_SEL_RANGE: (IndexError, "Port index is out of range"),                             # :269
    nat.ST_SEL_ORIGIN: (Exception, "Destination port must be different from current one"),    # :267
    nat.ST_NOT_AT_PORT: (Exception, "Not currently at port"),                                 # :343,352
    nat.ST_BAD_AMOUNT: (ValueError, "Invalid fuel amount"),                                   # :346,355
    nat.ST_UNKNOWN: (ValueError, "Action category unknown"),                                  # :374
    nat.ST_NO_PORTS: (Exception, "No ports available"),                                       # :360
    nat.ST_SAMPLE_CRASH: (TypeError, "'float' object is not subscriptable"),                  # :163
}

# byte offsets inside the per-instance 128-byte device / pinned-host block
_SHIP, _REWARD, _CLOCK, _ACTION, _DONE, _RESET, _ERR, _STATUS, _TAKEN, _WORDS, _NEXT, _BLOCK = 0, 16, 32, 48, 64, 68, 72, 76, 78, 80, 84, 128

_SHIP_FMT = struct.Struct("<BBBBid")            # shipenv_ship: x, y, origin, dest, cargo, fuel
_U32, _I16, _F64 = struct.Struct("<I"), struct.Struct("<h"), struct.Struct("<d")
_NEXT_FMT = struct.Struct("<hBB")                # step_sample: next action, words it used, its status
_FLAGS_FMT = struct.Struct("<IIIBxhB")           # done, reset, error words, status, taken, words used

# SHIPENV_NO_SERVER=1: every step() is one shipenv_step_sample launch + one stream synchronisation (A/B, debugging)
_USE_SERVER = os.environ.get("SHIPENV_NO_SERVER") != "1"

# The hot calls in C (csrc/facade_fast.c -> _facade_fast.so, built by __graft_entry__.build()): step() and the
# look-ahead sample_action() for the common shapes of a call; it declines everything else and the Python below runs,
# which stays the complete implementation.  SHIPENV_NO_FASTPATH=1 (or a missing .so) keeps it out.
_fast = None
_fast_tried = False


def _bind_fast():
    """Load and bind the C fast path once (needs libshipenv.so: its shipenv_serve_step is what the C code calls)."""
    global _fast, _fast_tried
    if _fast_tried:
        return
    _fast_tried = True
    if not _USE_SERVER or os.environ.get("SHIPENV_NO_FASTPATH") == "1":
        return
    try:
        import ctypes
        from . import _facade_fast as mod
        addr = ctypes.cast(nat.lib().shipenv_serve_step, ctypes.c_void_p).value
        mod.bind(addr, _ERRORS, nat.check, MOVES, nat.N_CARGO_ACTIONS, nat.N_FUEL_ACTIONS, nat.ACT_EXT_CARGO,
                 nat.ACT_EXT_FUEL)
        _fast = mod
    except ImportError:
        _fast = None


_contexts = {}   # device index -> (Context, world signature currently uploaded)


def _context(device_index):
    if device_index not in _contexts:
        _contexts[device_index] = [nat.Context(device_index), None]
    return _contexts[device_index]


class Environment:
    def __init__(self, map_path, game_size=(100, 100), seed=None, device=None, env_id=0):
        # ** Configurations (shipping/environment.py:29-43)
        self.size_game = game_size
        self.port_positions = []
        self.port_fuel = []
        self.port_cargo = []
        self.ship_position = []
        self.cargo = Initial.CARGO
        self.fuel = Initial.FUEL
        self.origin_port_index = None
        self.destination_port_index = None
        self.np_game = derive_grid(map_path, game_size)      # raises FileNotFoundError like :47

        if not torch.cuda.is_available():
            raise nat.ShipEnvError("Environment needs a CUDA device (sm_100a); there is no CPU fallback")
        self._device = torch.device(device if device is not None else "cuda")
        if self._device.index is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        _bind_fast()
        self._seed = _pyrandom.getrandbits(64) if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._env_id = int(env_id)
        self._clock = 0          # episode << 16 | step: selects the Philox cell of the next step
        self._word = 0           # words of that cell already consumed by sample_action()
        self._spec = None        # (record, clock, word, world) the last step's look-ahead sample_action was drawn for
        self._world_sig = None
        # One page-locked host block per instance; the N = 1 kernels read and write it in place through its
        # device-visible address (no staging copies: a step is one launch and one stream synchronisation).
        self._host = torch.zeros(_BLOCK, dtype=torch.uint8).pin_memory()
        self._hnp = self._host.numpy()
        self._buf = memoryview(self._hnp)
        self._host_ptr = self._host.data_ptr()
        self._base = nat.host_device_pointer(self._host_ptr)
        self._io = nat.StepIO()
        self._io.reward_f64, self._io.status = self._ptr(_REWARD), self._ptr(_STATUS)
        self._io.done_bits, self._io.reset_bits, self._io.error_bits = self._ptr(_DONE), self._ptr(_RESET), self._ptr(_ERR)
        self._passable_key = None
        self._passable = None
        self._ragged = False             # size_game is not the grid's shape (non-square game_size)
        self._sails = False              # set by _encode for a move outside ShipMove: its target is not ground
        self._w_game = self._w_ports = self._w_fuel = self._w_cargo = None   # what _sync_world last saw
        self._slot = None

    # ------------------------------------------------------------------ construction
    def _is_within_map(self, x, y):
        return 0 <= x < self.size_game[0] and 0 <= y < self.size_game[1]

    def add_port(self, pos):
        x, y = pos[0], pos[1]
        if not self._is_within_map(x, y):
            raise ValueError("Coordinates not within map")
        fuel, cargo = nat.port_stock(self._seed, len(self.port_positions))   # randint(5, 20) twice, :63-64
        self.port_positions.append(pos)
        self.port_fuel.append(fuel)
        self.port_cargo.append(cargo)
        self.np_game[x, y] = Entity.PORT

    def remove_port(self, idx):
        # the reference compares an int with a list here and always raises TypeError (:68)
        if 0 <= idx < self.port_positions and 0 <= idx < self.port_cargo:
            raise AssertionError("unreachable")
        raise IndexError("Invalid index")

    # ------------------------------------------------------------------ host <-> device
    def _sync_world(self):
        slot = _context(self._device.index)
        # fast path (every step of an agent's loop): the same grid object, the same port tables, and the context still
        # holds this world
        if (self.np_game is self._w_game and slot[1] is self._world_sig and self._world_sig is not None
                and self.port_positions == self._w_ports and self.port_fuel == self._w_fuel
                and self.port_cargo == self._w_cargo):
            return slot[0]
        # keyed on the array object itself (held, so its id cannot be recycled) and the port count: the passable set
        # only changes when np_game is replaced or add_port stamps a cell (it never shrinks)
        key = (self.np_game, len(self.port_positions))
        if self._passable_key is None or key[0] is not self._passable_key[0] or key[1] != self._passable_key[1]:
            g = np.asarray(self.np_game)
            if g.ndim != 2:
                raise ValueError("np_game must be a 2-D grid")
            # Non-square game_size: cv2.resize makes the grid size_game[1] x size_game[0] while _is_within_map bounds
            # x by size_game[0] and x indexes rows (:53, :99-100, :293).  The cells a ship can ever stand on are those
            # inside both, which is what the device gets as its map; a move that passes _is_within_map but leaves the
            # array raises numpy's IndexError in the reference, and in step() here before anything is launched.
            wd, hd = min(int(self.size_game[0]), int(g.shape[0])), min(int(self.size_game[1]), int(g.shape[1]))
            self._ragged = (wd, hd) != (int(self.size_game[0]), int(self.size_game[1]))
            self._passable = np.ascontiguousarray(g[:wd, :hd] != Entity.GROUND, dtype=np.uint8)
            self._passable_bytes = self._passable.tobytes()
            self._passable_key = key
        ports = [(int(p[0]), int(p[1])) for p in self.port_positions]
        sig = (self._passable_bytes, tuple(ports), tuple(map(int, self.port_fuel)),
               tuple(map(int, self.port_cargo)))
        if slot[1] != sig:
            if len(ports) > nat.MAX_PORTS:
                raise ValueError(f"at most {nat.MAX_PORTS} ports")
            slot[0].set_map(self._passable)
            slot[0].set_ports(ports, list(map(int, self.port_fuel)), list(map(int, self.port_cargo)))
            slot[1] = sig
        self._world_sig = slot[1]
        self._slot = slot                # (read by the C fast path, which makes the check above itself)
        self._w_game = self.np_game
        self._w_ports = [list(p) for p in self.port_positions]
        self._w_fuel, self._w_cargo = list(self.port_fuel), list(self.port_cargo)
        return slot[0]

    def _pack(self, action=0):
        pos = self.ship_position if len(self.ship_position) == 2 else (0, 0)
        origin, dest = self.origin_port_index, self.destination_port_index
        _SHIP_FMT.pack_into(self._buf, _SHIP, int(pos[0]) & 255, int(pos[1]) & 255,
                            nat.NONE_PORT if origin is None else int(origin),
                            nat.NONE_PORT if dest is None else int(dest), int(self.cargo), float(self.fuel))
        _U32.pack_into(self._buf, _CLOCK, self._clock)
        _I16.pack_into(self._buf, _ACTION, action)

    def _wait(self, ctx):
        """Returns once the launch has completed; the kernel's stores to the host block are then visible."""
        ctx.synchronize(self._stream())

    def _unpack(self):
        return _SHIP_FMT.unpack_from(self._buf, _SHIP)

    def _ptr(self, off):
        return self._base + off

    def _stream(self):
        return torch.cuda.current_stream(self._device).cuda_stream

    def _adopt(self, x, y, origin, dest, cargo, fuel, keep_int_fuel):
        self.origin_port_index = None if origin == nat.NONE_PORT else origin
        self.destination_port_index = None if dest == nat.NONE_PORT else dest
        self.cargo = cargo
        self.fuel = int(fuel) if keep_int_fuel and fuel == int(fuel) else fuel
        return x, y

    # ------------------------------------------------------------------ the reference API
    def _get_current_port_idx(self):
        for idx, port_position in enumerate(self.port_positions):
            if port_position == self.ship_position:
                return idx
        return None

    def _build_state(self):
        ship = {
            "position": self.ship_position,
            "fuel": self.fuel,
            "cargo": self.fuel,          # sic: the reference copies fuel here (:206); observation parity
            "origin_port_index": self.origin_port_index,
            "destination_port_index": self.destination_port_index,
        }
        ports = [{"index": idx, "position": p, "fuel": self.port_fuel[idx], "cargo": self.port_cargo[idx]}
                 for idx, p in enumerate(self.port_positions)]
        return {"ship": ship, "ports": ports}

    def reset(self):
        if len(self.port_positions) == 0:
            raise Exception("No ports available")                    # _sample_random_port, :156
        if len(self.port_positions) < 2:
            raise ValueError("reset needs at least two ports (the reference loops forever with one)")
        ctx = self._sync_world()
        self.np_game[(self.np_game == Entity.BOAT) | (self.np_game == Entity.TRAVEL)] = Entity.WATER
        self._pack()
        ctx.reset(self._ptr(_SHIP), self._ptr(_CLOCK), self._seed, self._env_id, 1, None, self._stream())
        self._wait(ctx)
        x, y, origin, dest, cargo, fuel = self._unpack()
        self._clock = _U32.unpack_from(self._buf, _CLOCK)[0]
        self._word = 0
        self._spec = None
        self._adopt(x, y, origin, dest, cargo, fuel, keep_int_fuel=True)
        current_port = self.port_positions[origin]
        self.np_game[current_port[0], current_port[1]] = Entity.BOAT
        self.ship_position = current_port                            # aliases the port's list, like :241
        return self._build_state()

    def _snapshot(self):
        """What the look-ahead draw of step() was made for: the attributes an agent could have changed since."""
        pos = self.ship_position
        return (pos[0], pos[1], self.origin_port_index, self.destination_port_index, self.cargo, self.fuel,
                self._clock, self._word, self._world_sig) if len(pos) == 2 else None

    def sample_action(self):
        if _fast is not None:
            r = _fast.sample_action(self)
            if r is not NotImplemented:
                return r
        ctx = self._sync_world()
        spec, self._spec = self._spec, None
        if spec is not None and spec == self._snapshot():
            # the previous step() already drew this call's action for exactly this state (shipenv_step_sample)
            taken, words, status = _NEXT_FMT.unpack_from(self._buf, _NEXT)
        else:
            self._pack()
            ctx.sample_action(self._ptr(_SHIP), self._ptr(_CLOCK), self._ptr(_TAKEN), self._ptr(_WORDS),
                              self._ptr(_STATUS), self._seed, self._env_id, 1, self._word, self._stream())
            self._wait(ctx)
            _, _, _, status, taken, words = _FLAGS_FMT.unpack_from(self._buf, _DONE)
        if status != nat.ST_OK:
            exc, msg = _ERRORS[status]
            raise exc(msg)
        self._word += words
        return self._decode(taken)

    def _decode(self, a):
        P = len(self.port_positions)
        if a < 4:
            return [ActionType.MOVE_SHIP, MOVES[a]]
        if a < 4 + P:
            return [ActionType.SELECT_PORT, a - 4]
        if a < 4 + P + nat.N_CARGO_ACTIONS:
            return [ActionType.TAKE_CARGO, a - (4 + P)]
        if a < 4 + P + nat.N_CARGO_ACTIONS + nat.N_FUEL_ACTIONS:
            return [ActionType.TAKE_FUEL, a - (4 + P + nat.N_CARGO_ACTIONS)]
        if a < nat.ACT_EXT_FUEL:
            return [ActionType.TAKE_CARGO, a - nat.ACT_EXT_CARGO]
        return [ActionType.TAKE_FUEL, a - nat.ACT_EXT_FUEL]

    def _encode(self, action):
        """[category, value] -> device action index; raises what the reference raises before any draw."""
        P = len(self.port_positions)
        category, value = action
        if category == ActionType.MOVE_SHIP:
            if len(value) != 2:
                raise ValueError("Move needs to be of format (x,y) or [x, y]")       # :274
            if self.destination_port_index is None:
                raise Exception("Cannot move without destination port")               # :276
            dx, dy = value
            if not (isinstance(dx, (int, np.integer)) and isinstance(dy, (int, np.integer))):
                self._float_move(dx, dy)
            key = (int(dx), int(dy))
            a = MOVE_INDEX.get(key)
            if a is not None:
                return a
            # any other (dx, dy): _move_ship takes it as it is (:278-287).  Its checks that come before the first draw
            # are made here: the bounds (:284) and -- for a non-square game_size -- numpy's own index check at :293.
            nx, ny = self.ship_position[0] + key[0], self.ship_position[1] + key[1]
            self._check_target(nx, ny)
            if max(abs(key[0]), abs(key[1])) > nat.MOVE_XY_MAX:
                raise NotImplementedError(f"moves of more than {nat.MOVE_XY_MAX} cells per axis have no action code")
            # a move of (0, 0) onto water "moves" too (:295-304): fuel becomes a float, the cell is repainted
            self._sails = self.np_game[nx, ny] != Entity.GROUND
            return nat.ACT_MOVE_XY_BASE - ((key[0] + 90) * 181 + (key[1] + 90))
        if category == ActionType.SELECT_PORT:
            if not 0 <= value < P:
                raise IndexError("Port index is out of range")                        # :269
            return 4 + int(value)
        if category in (ActionType.TAKE_CARGO, ActionType.TAKE_FUEL):
            if value != int(value):
                raise NotImplementedError("fractional amounts are not supported on the device")
            v = min(max(int(value), 0), 8191)      # anything outside 1..255 is rejected by the kernel as well
            if category == ActionType.TAKE_CARGO:
                return 4 + P + v if v < nat.N_CARGO_ACTIONS else nat.ACT_EXT_CARGO + v
            return 4 + P + nat.N_CARGO_ACTIONS + v if v < nat.N_FUEL_ACTIONS else nat.ACT_EXT_FUEL + v
        raise ValueError("Action category unknown")                                   # :374

    def _check_target(self, nx, ny):
        """What _move_ship checks about the target cell before its first draw (:284) and, when game_size is not the
        shape of the grid cv2.resize made (:53), numpy's IndexError at :293 -- raised here before anything is launched
        (the reference has drawn the fuel cost by then, but no state has changed and the cell starts over)."""
        if not self._is_within_map(nx, ny):
            raise ValueError("Move is out of range")                                  # :284
        rows, cols = np.shape(self.np_game)
        if nx >= rows:
            raise IndexError(f"index {nx} is out of bounds for axis 0 with size {rows}")
        if ny >= cols:
            raise IndexError(f"index {ny} is out of bounds for axis 1 with size {cols}")

    def _float_move(self, dx, dy):
        """A move with a non-integer component: the reference gets as far as indexing np_game with a float (:293)."""
        nx, ny = self.ship_position[0] + dx, self.ship_position[1] + dy
        if not self._is_within_map(nx, ny):
            raise ValueError("Move is out of range")                                  # :284
        raise IndexError("only integers, slices (`:`), ellipsis (`...`), numpy.newaxis (`None`) and integer or "
                         "boolean arrays are valid indices")

    def step(self, action):
        if _fast is not None:
            r = _fast.step(self, action)
            if r is not NotImplemented:
                return r
        if len(self.port_positions) == 0:
            raise Exception("No ports available")                                     # :360
        a = self._encode(action)
        ctx = self._sync_world()
        old = (int(self.ship_position[0]), int(self.ship_position[1])) if len(self.ship_position) == 2 else None
        fuel_was_int = isinstance(self.fuel, int)
        free_move = a <= nat.ACT_MOVE_XY_BASE
        if self._ragged and not free_move and action[0] == ActionType.MOVE_SHIP and old is not None:
            dx, dy = MOVES[a]
            if self._is_within_map(old[0] + dx, old[1] + dy):      # off-map moves are the device's to reject
                self._check_target(old[0] + dx, old[1] + dy)
        self._pack(a)
        # the step and, with it, the draw of the sample_action() call that usually follows it
        if _USE_SERVER:
            # answered by the resident server kernel of this device (no launch, no stream synchronisation)
            ctx.serve_step(self._host_ptr, self._seed, self._env_id, 0, 0, self._word)
        else:
            ctx.step_sample(self._ptr(_SHIP), self._ptr(_CLOCK), self._ptr(_ACTION), self._io, self._ptr(_NEXT),
                            self._ptr(_NEXT + 2), self._ptr(_NEXT + 3), self._seed, self._env_id, 1, 0, 0, self._word,
                            self._stream())
            self._wait(ctx)
        x, y, origin, dest, cargo, fuel = self._unpack()
        done_word, _, _, status, _, _ = _FLAGS_FMT.unpack_from(self._buf, _DONE)
        if status != nat.ST_OK:       # nothing changed on the device, no draw was consumed: the cell continues
            self._spec = self._snapshot()
            exc, msg = _ERRORS[status]
            raise exc(msg)
        self._clock = _U32.unpack_from(self._buf, _CLOCK)[0]
        self._word = 0
        moved = self._sails if free_move else (x, y) != old
        keep_int = fuel_was_int and not (action[0] == ActionType.MOVE_SHIP and moved)
        self._adopt(x, y, origin, dest, cargo, fuel, keep_int_fuel=keep_int)
        if moved:
            self.ship_position = [x, y]                                               # a fresh list, like :297
            self.np_game[old[0], old[1]] = Entity.TRAVEL                              # :303-304
            self.np_game[x, y] = Entity.BOAT
        self._spec = self._snapshot()
        reward = _F64.unpack_from(self._buf, _REWARD)[0]
        return self._build_state(), reward, bool(done_word & 1), {}

    # ------------------------------------------------------------------ extensions: MCTS on the device
    def _root_ptr(self):
        self._pack()
        return self._ptr(_SHIP)      # the one root record is read from the host block in place

    def rollout_returns(self, simulations=50, max_rollout_steps=100, root_id=None):
        """``MCTSAgent._rollout`` (agents/mcts.py:211-238) x ``simulations`` from the current state in ONE launch
        (the reference builds an env copy and loops in Python for each).  Returns a list of floats."""
        ctx = self._sync_world()
        roots = self._root_ptr()
        ret = torch.empty(simulations, dtype=torch.float64, device=self._device)
        ctx.rollout(roots, 1, simulations, max_rollout_steps, ret.data_ptr(), None, None, None,
                    self._seed, self._clock if root_id is None else root_id, self._stream())
        return ret.cpu().tolist()

    def mcts_action(self, simulations=50, max_rollout_steps=100, c_param=1.4, root_id=None):
        """``MCTSAgent.choose_action`` (agents/mcts.py:240-292) from the current state in ONE launch: UCB1 tree,
        expansion, rollouts, backpropagation and the final pick all run on the device.  Returns the action in the
        environment's format, ``[ActionType.MOVE_SHIP, [dx, dy]]``; falls back to ``sample_action()`` when the tree
        has no children or the state is not the four-move case, as the reference does (:291-292)."""
        ctx = self._sync_world()
        roots = self._root_ptr()
        dev = self._device
        act = torch.empty(1, dtype=torch.int16, device=dev)
        visits = torch.empty(1, dtype=torch.int32, device=dev)
        ca = torch.empty(4, dtype=torch.int16, device=dev)
        cv = torch.empty(4, dtype=torch.int32, device=dev)
        ct = torch.empty(4, dtype=torch.float64, device=dev)
        st = torch.empty(1, dtype=torch.uint8, device=dev)
        log_table = np.concatenate([[0.0], np.log(np.arange(1, simulations + 2))])
        ctx.mcts(roots, 1, simulations, max_rollout_steps, c_param, log_table.ctypes.data, act.data_ptr(),
                 visits.data_ptr(), ca.data_ptr(), cv.data_ptr(), ct.data_ptr(), st.data_ptr(), self._seed,
                 self._clock if root_id is None else root_id, self._stream())
        a = int(act.item())
        if a < 0:
            return self.sample_action()
        return [ActionType.MOVE_SHIP, list(MOVES[a])]

    # ------------------------------------------------------------------ rendering (host, cv2; not on the hot path)
    _PALETTE = ((Entity.WATER, Color.WATER), (Entity.GROUND, Color.GROUND), (Entity.PORT, Color.PORT),
                (Entity.DESTINY_PORT, Color.DESTINY_PORT), (Entity.TRAVEL, Color.TRAVEL))
    # the boat sprite of the reference renderer as (row offsets, column offsets, colour)
    _SPRITE = (((0, 1), (-3, 3), Color.BOAT), ((1, 2), (-2, 2), Color.BOAT), ((-4, 0), (0, 1), Color.BLACK),
               ((-4, -1), (1, 2), Color.WHITE), ((-3, -2), (2, 3), Color.WHITE))

    def render_real_time(self, render_size=(355, 533)):
        """Same picture as the reference's renderer (shipping/environment.py:106-143), drawn on the host."""
        import cv2
        grid = self.np_game
        canvas = np.zeros(grid.shape + (3,), dtype=np.uint8)
        for code, bgr in self._PALETTE:
            canvas[grid == code] = bgr
        boats = np.argwhere(grid == Entity.BOAT)
        if len(boats):
            bx, by = (int(v) for v in boats[0])
            for (r0, r1), (c0, c1), bgr in self._SPRITE:
                canvas[bx + r0:bx + r1, by + c0:by + c1] = bgr
        frame = cv2.resize(canvas, render_size, interpolation=cv2.INTER_AREA)
        lines = [(20, f"Fuel: {self.fuel:.2f}"), (40, f"Cargo: {self.cargo}")]
        lines += [(80 + 20 * k, f"P{k + 1} - Cargo: {c} - Fuel: {f}")
                  for k, (c, f) in enumerate(zip(self.port_cargo, self.port_fuel))]
        for ypix, text in lines:
            cv2.putText(frame, text, (10, ypix), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 255, 255), 1, cv2.LINE_AA)
        cv2.imshow("Environment", frame)
        if cv2.waitKey(100) == ord("q"):
            cv2.destroyAllWindows()
            raise SystemExit
