"""CPU (-m "not gpu"): the C-ABI library loads and exports exactly what include/shipenv.h declares, host
logic (map pipeline, action encoding, sharding), the product never touches the oracle, and the product
fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden

HEADER = os.path.join(ROOT, "include", "shipenv.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shipenv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from shippingenv_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = header_functions()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/shipenv.h but not exported"
    assert set(names) == set(_native._SIGNATURES), "ctypes binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (shipenv_[a-z0-9_]+)", out))
    assert exported == set(names), exported ^ set(names)


def test_library_is_sm100a_with_tma():
    from shippingenv_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass        # TMA bulk copies + mbarriers in the step kernels


def test_struct_layouts_match_header():
    from shippingenv_b200 import _native
    assert ctypes.sizeof(_native.Consts) == 8 + 8 + 11 * 8
    assert ctypes.sizeof(_native.StepIO) == 11 * 8
    assert _native.lib().shipenv_abi_version() == 2


def test_host_philox_and_port_stock_without_gpu():
    from shippingenv_b200 import _native
    assert _native.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    z = golden("port_stock")
    for seed in (0x5EED, 0, 0xDEADBEEFCAFEF00D):
        got = [_native.port_stock(seed, k) for k in range(5)]
        assert [g[0] for g in got] == z[f"fuel_{seed:x}"].tolist()
        assert [g[1] for g in got] == z[f"cargo_{seed:x}"].tolist()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shippingenv_b200 import BatchedEnvironment, Environment, ShipEnvError
    with pytest.raises(ShipEnvError):
        BatchedEnvironment(4, passable=golden("default_map")["passable"])
    with pytest.raises(ShipEnvError):
        Environment(os.path.join(GOLDEN, "default_map.npz"))
    from shippingenv_b200 import _native
    h = ctypes.c_void_p()
    assert _native.lib().shipenv_ctx_create(0, ctypes.byref(h)) != 0
    assert _native.lib().shipenv_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shippingenv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
    code = ("import sys; import shippingenv_b200, shippingenv_b200.environment, shippingenv_b200.preprocessing;"
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_map_pipeline_matches_cv2_reference_steps(tmp_path):
    """derive_grid == imread(gray) -> [50:200, 100:300] -> resize INTER_AREA -> threshold 128 (env:45-55)."""
    import cv2
    from shippingenv_b200.mapio import derive_grid, passable_of
    rng = np.random.default_rng(3)
    img = (rng.random((355, 533)) * 255).astype(np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), 6)
    img = ((img.astype(np.float32) - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
    path = str(tmp_path / "map.png")
    cv2.imwrite(path, img)
    for size in ((100, 100), (64, 64)):
        want = cv2.imread(path, cv2.IMREAD_GRAYSCALE)[50:200, 100:300]
        want = cv2.resize(want, size, interpolation=cv2.INTER_AREA)
        _, want = cv2.threshold(want, 128, 1, cv2.THRESH_BINARY)
        got = derive_grid(path, size)
        assert got.dtype == int and (got == want).all() and 0 < got.sum() < got.size
        assert (passable_of(got) == (got != 0)).all()
    with pytest.raises(FileNotFoundError):
        derive_grid(str(tmp_path / "missing.jpg"))
    g = derive_grid(os.path.join(GOLDEN, "default_map.npz"))
    assert g.shape == (100, 100) and g.sum() == 5709


def test_preprocessing_helpers():
    from shippingenv_b200.preprocessing import (preprocess_state, discretize_state, map_action_to_env_action,
                                                get_action_space_size, find_current_port_index)
    from shippingenv_b200 import ActionType, ShipMove

    class E:
        port_positions = [[1, 1], [2, 2], [3, 3], [4, 4], [5, 5]]
    state = {"ship": {"position": [2, 2], "fuel": 150.5, "cargo": 150.5, "origin_port_index": None,
                      "destination_port_index": 3},
             "ports": [{"index": k, "position": p, "fuel": 10 + k, "cargo": 5 + k} for k, p in enumerate(E.port_positions)]}
    row = preprocess_state(state)
    assert row.shape == (1, 26) and row[0, :6].tolist() == [2, 2, 150.5, 150.5, -1, 3]
    assert row[0, 6:10].tolist() == [1, 1, 10, 5]
    assert discretize_state(state) == ((2, 2), 4, 4, -1, 3)
    assert find_current_port_index(state) == 1
    assert get_action_space_size(E) == 259
    assert map_action_to_env_action(0, E) == [ActionType.MOVE_SHIP, ShipMove.NORTH]
    assert map_action_to_env_action(1, E) == [ActionType.MOVE_SHIP, ShipMove.EAST]
    assert map_action_to_env_action(3, E) == [ActionType.MOVE_SHIP, ShipMove.WEST]
    assert map_action_to_env_action(8, E) == [ActionType.SELECT_PORT, 4]
    assert map_action_to_env_action(9, E) == [ActionType.TAKE_CARGO, 0]
    assert map_action_to_env_action(58, E) == [ActionType.TAKE_CARGO, 49]
    assert map_action_to_env_action(59, E) == [ActionType.TAKE_FUEL, 0]
    assert map_action_to_env_action(258, E) == [ActionType.TAKE_FUEL, 199]


def test_action_codes_match_the_header():
    """The action-space constants of the Python layer are the header's, and the move code of any (dx, dy) is the one the
    reference harness uses (oracle/ref_harness.py: move_xy) -- the four moves of ShipMove keep their flat indices."""
    import re
    from shippingenv_b200 import _native
    from shippingenv_b200.preprocessing import move_action
    from oracle.ref_harness import move_xy
    hdr = open(os.path.join(ROOT, "include", "shipenv.h")).read()
    macro = {m.group(1): m.group(2) for m in re.finditer(r"#define (SHIPENV_\w+) \(?(-?\w+)\)?", hdr)}
    assert int(macro["SHIPENV_ACT_POLICY"]) == _native.ACT_POLICY == -1
    assert int(macro["SHIPENV_ACT_NONE"]) == _native.ACT_NONE == -32768
    assert int(macro["SHIPENV_ACT_MOVE_XY_BASE"]) == _native.ACT_MOVE_XY_BASE == -2
    assert int(macro["SHIPENV_MOVE_XY_MAX"]) == _native.MOVE_XY_MAX == 90
    assert int(macro["SHIPENV_ACT_EXT_CARGO"], 16) == _native.ACT_EXT_CARGO
    assert int(macro["SHIPENV_ACT_EXT_FUEL"], 16) == _native.ACT_EXT_FUEL
    assert [move_action(*m) for m in ((0, -1), (-1, 0), (0, 1), (1, 0))] == [0, 1, 2, 3]
    seen = set()
    for dx in (-90, -7, -1, 0, 1, 2, 90):
        for dy in (-90, -3, -1, 0, 1, 5, 90):
            if abs(dx) + abs(dy) == 1:
                continue
            a = move_action(dx, dy)
            assert a == move_xy(dx, dy) and -32762 <= a <= -2 and a not in seen
            seen.add(a)
    assert move_action(0, 0) == -2 - (90 * 181 + 90) and move_action(90, 90) == -32762 and move_action(-90, -90) == -2
    with pytest.raises(ValueError):
        move_action(91, 0)


def test_shard_ranges_partition_the_env_ids():
    from shippingenv_b200.stats import shard_range
    for total in (16, 17, 1 << 24, 5):
        for ws in (1, 2, 4, 8):
            edges = [shard_range(total, r, ws) for r in range(ws)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert shard_range(1 << 24, 3, 8) == (3 << 21, 4 << 21)


def test_sharding_invariance_on_the_oracle():
    """Global env ids make a trajectory independent of how envs are split over ranks (checked on the CPU
    oracle here; tests/test_gpu_parity.py checks the kernels against the same oracle)."""
    from oracle import lib as orc
    from shippingenv_b200.stats import shard_range
    z = golden("default_map")
    o = orc.Oracle(z["passable"], z["ports"], z["stock_fuel"], z["stock_cargo"])
    n, seed, T = 1000, 77, 40
    rng = np.random.default_rng(5)
    acts = rng.integers(0, 4, (T, n)).astype(np.int16)
    whole, wclock = np.zeros(n, orc.SHIP_DTYPE), np.zeros(n, np.uint32)
    o.reset(whole, wclock, seed)
    for t in range(T):
        o.step(whole, wclock, acts[t], seed, 0, 25, True)
    for ws in (2, 8):
        parts = []
        for r in range(ws):
            lo, hi = shard_range(n, r, ws)
            s, c = np.zeros(hi - lo, orc.SHIP_DTYPE), np.zeros(hi - lo, np.uint32)
            o.reset(s, c, seed, env_base=lo)
            for t in range(T):
                o.step(s, c, acts[t, lo:hi], seed, lo, 25, True)
            parts.append(s)
        assert np.concatenate(parts).tobytes() == whole.tobytes()
