"""CPU study for the next round's GP column solve: how much precision does V = L^-1 K*^T need?

The observation only uses the solve for float32 features: the error means  v^T z  and the normalised variance
1 - |v|^2 / sigma^2  (env/wind_gp.py:143-207, env/features.py:430-520), with 1e-4 absolute of head-room
(SURVEY.md section 8 row f1).  `k_gp_column4` runs the solve in fp64 on the DMMA pipe at about half of its
bound; this script emulates cheaper arithmetic in NumPy on synthetic 120-measurement windows that look like
a flight (3-minute spacing, ~10 m/s drift, piecewise-constant pressure with ACS moves) and reports the
worst feature error of each variant against fp64:

  fp32        every operand and accumulator float32 (FFMA)
  tf32        operands rounded to TF32 (10-bit mantissa), float32 accumulation (one tensor-core pass)
  tf32x3      operands split hi + lo, three TF32 products (hi*hi + hi*lo + lo*hi), float32 accumulation
  bf16x3      the same with bf16 splits (8-bit mantissa), three terms

The factor L itself stays fp64 -> rounded once to the working precision (k_gp_update is a different kernel).
Stand-alone (NumPy only; no oracle, no GPU):   python scripts/gp_precision_study.py > profiles/r01c_gp_precision_study.json
"""
import json

import numpy as np

LENGTH = np.array([357000.0, 357000.0, 326.0, 34560.0])      # env/wind_gp.py:33-35
SIGMA2, NOISE = 3.6 ** 2, 0.05
LEVELS = np.linspace(5000.0, 14000.0, 181)
BLOCK = 8


def kernel(a, b):
  d = (a[:, None, :] - b[None, :, :]) / LENGTH
  return SIGMA2 * np.exp(-np.sqrt((d * d).sum(-1)))


def flight(rng, m=120):
  """m measurements, 180 s apart: wind drift + pressure plateaus joined by ACS ramps."""
  t = np.arange(m) * 180.0
  vel = rng.normal(0, 8.0, 2)
  xy = np.cumsum(rng.normal(vel, 2.0, (m, 2)) * 180.0, axis=0)
  p = np.empty(m); level = rng.uniform(6000, 13000); target = level
  for i in range(m):
    if rng.random() < 0.04:
      target = rng.uniform(6000, 13000)
    level += np.clip(target - level, -60.0, 60.0)            # ~60 Pa per agent step while the ACS runs
    p[i] = level + rng.normal(0, 1.5)
  err = rng.normal(0, 2.0, (m, 2))                            # forecast errors, m/s
  return np.column_stack([xy, p, t]), err


def round_mantissa(x, bits):
  """float32 value rounded (to nearest) to `bits` explicit mantissa bits."""
  x = np.asarray(x, np.float32)
  u = x.view(np.uint32).astype(np.uint64)
  drop = 23 - bits
  u = (u + (1 << (drop - 1))) & ~np.uint64((1 << drop) - 1)
  return u.astype(np.uint32).view(np.float32)


def matmul_variant(a, b, variant):
  """a @ b with the operand rounding of `variant`, float32 accumulation."""
  a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
  if variant == 'fp32':
    return a @ b
  bits = 10 if variant.startswith('tf32') else 7
  a_hi, b_hi = round_mantissa(a, bits), round_mantissa(b, bits)
  if not variant.endswith('x3'):
    return a_hi @ b_hi
  a_lo, b_lo = round_mantissa(a - a_hi, bits), round_mantissa(b - b_hi, bits)
  return a_hi @ b_hi + (a_hi @ b_lo + a_lo @ b_hi)


def blocked_solve(l64, rhs64, variant):
  """Right-looking blocked forward substitution as k_gp_column4 does it (8 x 8 blocks, inverted diagonal
  blocks prepared in fp64), with the block products in `variant` arithmetic."""
  m = l64.shape[0]
  nb = m // BLOCK
  c = rhs64.astype(np.float32).copy()
  v = np.zeros_like(c)
  for j in range(nb):
    js = slice(j * BLOCK, (j + 1) * BLOCK)
    inv_diag = np.linalg.inv(l64[js, js])
    v[js] = matmul_variant(inv_diag, c[js], variant)
    if j + 1 < nb:
      rest = slice((j + 1) * BLOCK, m)
      c[rest] -= matmul_variant(l64[rest, js], v[js], variant)
  return v


def main():
  rng = np.random.default_rng(0)
  variants = ['fp32', 'tf32', 'tf32x3', 'bf16x3']
  worst = {v: {'variance_abs': 0.0, 'mean_abs_mps': 0.0} for v in variants}
  conds = []
  n_windows = 200
  for _ in range(n_windows):
    loc, err = flight(rng)
    k = kernel(loc, loc) + NOISE * np.eye(len(loc))
    conds.append(float(np.linalg.cond(k)))
    l = np.linalg.cholesky(k)
    z = np.linalg.solve(l, err)                                               # fp64, from k_gp_update
    q = np.column_stack([np.full(181, loc[-1, 0]), np.full(181, loc[-1, 1]), LEVELS, np.full(181, loc[-1, 3])])
    k_star = kernel(loc, q)                                                   # [m, 181]
    v64 = np.linalg.solve(l, k_star)
    var64 = np.maximum(SIGMA2 - (v64 ** 2).sum(0), 0.0) / SIGMA2
    mean64 = v64.T @ z
    for name in variants:
      v = blocked_solve(l, k_star, name)
      var = np.maximum(np.float32(SIGMA2) - (v * v).sum(0, dtype=np.float32), 0.0) / np.float32(SIGMA2)
      mean = v.T @ z.astype(np.float32)
      w = worst[name]
      w['variance_abs'] = max(w['variance_abs'], float(np.abs(var - var64).max()))
      w['mean_abs_mps'] = max(w['mean_abs_mps'], float(np.abs(mean - mean64).max()))
  # the wind-magnitude feature squashes m/s by x / (x + 30): d(feature) <= d(m/s) / 30
  for w in worst.values():
    w['mean_feature_abs_bound'] = w['mean_abs_mps'] / 30.0
  print(json.dumps({'windows': n_windows, 'measurements': 120, 'levels': 181,
                    'cond_K_median': float(np.median(conds)), 'cond_K_max': float(np.max(conds)),
                    'tolerance_abs': 1e-4, 'worst_error_vs_fp64': worst}, indent=1))


if __name__ == '__main__':
  main()
