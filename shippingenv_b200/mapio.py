"""Map construction on the host (the JPEG never goes to the GPU, only the derived grid).

``derive_grid`` is the reference's ``Environment._initialize_map``
(shipping/environment.py:45-55): gray decode -> crop [50:200, 100:300] -> INTER_AREA resize
to ``game_size`` -> threshold at 128 -> {0 ground, 1 water}.  The decode is cv2's, on
purpose: 94 of the 10,000 default cells sit within +-8 of the threshold after the resize, so
a re-implemented resampler would not reproduce the map.

A pre-derived grid can be given instead of an image: ``.npy`` (the int grid) or ``.npz``
with an ``np_game`` entry (``tests/golden/default_map.npz`` is the reference's default map).
"""
import os

import numpy as np

_CROP_X = (100, 300)   # shipping/environment.py:49
_CROP_Y = (50, 200)    # shipping/environment.py:50
_cache = {}


def derive_grid(map_path, game_size=(100, 100)):
    """Return the int64 ``np_game`` grid of a fresh reference ``Environment`` (before ports)."""
    key = (os.path.abspath(map_path), tuple(game_size))
    try:
        stamp = os.path.getmtime(map_path)
    except OSError:
        raise FileNotFoundError(f"Cannot read the image at {map_path}")  # shipping/environment.py:47
    hit = _cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1].copy()
    ext = os.path.splitext(map_path)[1].lower()
    if ext == ".npy":
        grid = np.load(map_path)
    elif ext == ".npz":
        grid = np.load(map_path)["np_game"]
    else:
        import cv2
        img = cv2.imread(map_path, cv2.IMREAD_GRAYSCALE)
        if img is None:
            raise FileNotFoundError(f"Cannot read the image at {map_path}")
        img = img[_CROP_Y[0]:_CROP_Y[1], _CROP_X[0]:_CROP_X[1]]
        img = cv2.resize(img, tuple(game_size), interpolation=cv2.INTER_AREA)
        _, grid = cv2.threshold(img, 128, 1, cv2.THRESH_BINARY)
    grid = np.asarray(grid).astype(int)
    # cv2.resize(img, (w, h)) returns h rows and w columns (:53): that is the shape a stored grid must have too
    if ext in (".npy", ".npz") and tuple(grid.shape) != (game_size[1], game_size[0]):
        raise ValueError(f"grid in {map_path} has shape {grid.shape}, expected {(game_size[1], game_size[0])}")
    _cache[key] = (stamp, grid)
    return grid.copy()


def passable_of(np_game):
    """Cells a ship can enter: everything that is not Entity.GROUND (shipping/environment.py:293)."""
    return (np.asarray(np_game) != 0).astype(np.uint8)
