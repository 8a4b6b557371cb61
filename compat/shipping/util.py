"""shipping/util.py:3-10 helpers (host-side; the kernels use integer forms of both)."""
import numpy as np


def calculate_euclidean_distance(start, end):
    return np.sqrt((end[0] - start[0]) ** 2 + (end[1] - start[1]) ** 2)


def normalize(raw, max, min):
    return raw / max if min == 0 else (raw - min) / (max - min)
