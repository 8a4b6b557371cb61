"""Host-side helpers with the names of shipping/util.py (the kernels use integer forms of both)."""
import math


def calculate_euclidean_distance(start, end):
    """Distance between two [x, y] cells (shipping/util.py:3-4)."""
    return math.hypot(end[0] - start[0], end[1] - start[1])


def normalize(raw, max, min):
    """Position of ``raw`` in [min, max] as a fraction (shipping/util.py:6-10)."""
    span = max - min
    return (raw - min) / span
