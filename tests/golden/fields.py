"""Deterministic synthetic wind fields shared by the golden generator and the tests.

Values come from small-integer arithmetic scaled by powers of two, so they are
bit-identical fp32 on every machine (no libm, no RNG stream dependence).
Shape [21, 21, 10, 9, 2] = (x, y, pressure, time, uv) as in generative/vae.py:38-51.
"""
import numpy as np

SHAPE = (21, 21, 10, 9, 2)


def smooth_field(variant: int = 0) -> np.ndarray:
  """A smooth, sheared field with |u|,|v| up to ~12 m/s (realistic magnitudes)."""
  i, j, k, l = np.meshgrid(np.arange(21), np.arange(21), np.arange(10), np.arange(9), indexing='ij')
  a = variant
  u = ((i - 10) * (3 + a) + (k - 5) * 7 - l * 2 + (j - 10) * (k - 4)) / 8.0
  v = ((j - 10) * 5 - (k - 4 - a) * (l - 4) * 2 + (i - 10) * (2 - k)) / 8.0
  return np.stack([u, v], axis=-1).astype(np.float32)


def hashed_field(seed: int = 0) -> np.ndarray:
  """A rough field: per-node integer hash mapped to [-16, 16) with 2^-11 resolution."""
  idx = np.arange(np.prod(SHAPE), dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
  z = idx
  z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
  z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
  z = z ^ (z >> np.uint64(31))
  vals = ((z >> np.uint64(48)).astype(np.int64) - 32768).astype(np.float64) / 2048.0
  return vals.reshape(SHAPE).astype(np.float32)


def field_bank() -> np.ndarray:
  """[4, 21, 21, 10, 9, 2]: two smooth variants + two hashed fields."""
  return np.stack([smooth_field(0), smooth_field(1), hashed_field(0), hashed_field(1)])
