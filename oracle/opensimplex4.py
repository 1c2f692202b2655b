"""4-D OpenSimplex noise, restated from the published algorithm (ORACLE / test infrastructure).

PARITY UNPINNED.  The reference calls the third-party package
`opensimplex==0.3` (requirements.txt:43; call sites
env/simplex_wind_noise.py:107 `OpenSimplex(seed=...)` and :142-146
`noise4d(x, y, z, w)`).  That package is neither vendored under
/root/reference nor installed in this image, and the reference's own tests
only assert determinism and forecast != ground truth
(env/grid_based_wind_field_test.py:67-84).  This file restates the
published OpenSimplex (K. Spencer, 2014) definition that package implements:

  * permutation table: 256 entries shuffled by a 64-bit LCG
    (x <- x * 6364136223846793005 + 1442695040888963407, wrapped to int64),
    three warm-up rounds, then a Fisher-Yates pass from i = 255 down to 0
    with r = (seed + 31) mod (i + 1);
  * lattice: Z^4 in "stretched" space, xs = x + STRETCH * (x+y+z+w), with
    STRETCH = (1/sqrt(5) - 1)/4 and the inverse SQUISH = (sqrt(5) - 1)/4;
  * each lattice vertex v contributes  max(0, 2 - |d|^2)^4 * (g(v) . d),
    d = displacement from the un-stretched vertex, g(v) one of 64 gradients
    (permutations of (+-3, +-1, +-1, +-1)) selected by
    perm[(perm[(perm[(perm[x&255]+y)&255]+z)&255]+w)&255] & 0xFC;
  * result = sum / 30.

The package finds contributing vertices with a large hand-unrolled decision
tree; here the sum runs over EVERY vertex whose kernel is non-zero (at most
one coordinate can leave the unit hypercube around the point, so 16 + 64
candidates are tested).  The two agree wherever the decision tree visits all
in-range vertices.  The CUDA kernel implements exactly this definition, so
oracle <-> GPU parity is exact; reference <-> oracle parity for noise VALUES
is unpinned (only the variance constant OPENSIMPLEX_VARIANCE = 0.0569,
env/simplex_wind_noise.py:69, can be checked statistically).
"""
import numpy as np

STRETCH_4D = -0.138196601125011   # (1/sqrt(4+1)-1)/4
SQUISH_4D = 0.309016994374947     # (sqrt(4+1)-1)/4
NORM_4D = 30.0

_M64 = (1 << 64) - 1


def _build_gradients():
  g = np.zeros((64, 4), np.float64)
  for r in range(16):
    for j in range(4):
      v = [1.0, 1.0, 1.0, 1.0]
      v[j] = 3.0
      for c in range(4):
        if (r >> c) & 1:
          v[c] = -v[c]
      g[r * 4 + j] = v
  return g


GRADIENTS_4D = _build_gradients()   # index = (perm value & 0xFC) >> 2


def _wrap_i64(v: int) -> int:
  v &= _M64
  return v - (1 << 64) if v >= (1 << 63) else v


def make_perm(seed: int) -> np.ndarray:
  """256-entry permutation table for `seed` (OpenSimplex.__init__)."""
  perm = np.zeros(256, np.uint8)
  source = list(range(256))
  s = int(seed)
  for _ in range(3):
    s = _wrap_i64(s * 6364136223846793005 + 1442695040888963407)
  for i in range(255, -1, -1):
    s = _wrap_i64(s * 6364136223846793005 + 1442695040888963407)
    r = (s + 31) % (i + 1)      # Python modulo: already non-negative
    perm[i] = source[r]
    source[r] = source[i]
  return perm


def make_perms(seeds) -> np.ndarray:
  """Vectorised make_perm: seeds[...]-> uint8[..., 256]."""
  seeds = np.asarray(seeds)
  flat = seeds.reshape(-1)
  out = np.empty((flat.size, 256), np.uint8)
  for k, s in enumerate(flat):
    out[k] = make_perm(int(s))
  return out.reshape(seeds.shape + (256,))


# Candidate lattice offsets relative to floor(stretched point): the 16 cube
# corners, then for each corner and axis the neighbour one step further out.
def _candidates():
  out = []
  for m in range(16):
    base = [(m >> c) & 1 for c in range(4)]
    out.append(tuple(base))
    for c in range(4):
      v = list(base)
      v[c] = 2 if base[c] == 1 else -1
      out.append(tuple(v))
  return out


CANDIDATES = _candidates()      # 80 offsets


def noise4d_scalar(perm, x, y, z, w) -> float:
  s = (x + y + z + w) * STRETCH_4D
  xs, ys, zs, ws = x + s, y + s, z + s, w + s
  xb, yb, zb, wb = (int(np.floor(v)) for v in (xs, ys, zs, ws))
  q = (xb + yb + zb + wb) * SQUISH_4D
  dx0, dy0, dz0, dw0 = x - (xb + q), y - (yb + q), z - (zb + q), w - (wb + q)
  value = 0.0
  for (i, j, k, l) in CANDIDATES:
    t = (i + j + k + l) * SQUISH_4D
    dx, dy, dz, dw = dx0 - i - t, dy0 - j - t, dz0 - k - t, dw0 - l - t
    attn = 2.0 - dx * dx - dy * dy - dz * dz - dw * dw
    if attn > 0.0:
      h = int(perm[(int(perm[(int(perm[(int(perm[(xb + i) & 255]) + yb + j) & 255]) + zb + k) & 255])
                    + wb + l) & 255])
      g = GRADIENTS_4D[h >> 2]
      attn *= attn
      value += attn * attn * (g[0] * dx + g[1] * dy + g[2] * dz + g[3] * dw)
  return value / NORM_4D


def noise4d(perm, x, y, z, w) -> np.ndarray:
  """Vectorised noise: perm uint8[..., 256] broadcast against x,y,z,w[...]."""
  x, y, z, w = (np.asarray(v, np.float64) for v in (x, y, z, w))
  shape = np.broadcast_shapes(x.shape, y.shape, z.shape, w.shape, perm.shape[:-1])
  x, y, z, w = (np.broadcast_to(v, shape).reshape(-1) for v in (x, y, z, w))
  pm = np.broadcast_to(perm, shape + (256,)).reshape(-1, 256)
  n = x.size
  rows = np.arange(n)
  s = (x + y + z + w) * STRETCH_4D
  xs, ys, zs, ws = x + s, y + s, z + s, w + s
  xb = np.floor(xs).astype(np.int64); yb = np.floor(ys).astype(np.int64)
  zb = np.floor(zs).astype(np.int64); wb = np.floor(ws).astype(np.int64)
  q = (xb + yb + zb + wb) * SQUISH_4D
  dx0, dy0, dz0, dw0 = x - (xb + q), y - (yb + q), z - (zb + q), w - (wb + q)
  value = np.zeros(n)
  for (i, j, k, l) in CANDIDATES:
    t = (i + j + k + l) * SQUISH_4D
    dx, dy, dz, dw = dx0 - i - t, dy0 - j - t, dz0 - k - t, dw0 - l - t
    attn = 2.0 - dx * dx - dy * dy - dz * dz - dw * dw
    m = attn > 0.0
    if not m.any():
      continue
    r = rows[m]
    h = pm[r, (xb[m] + i) & 255].astype(np.int64)
    h = pm[r, (h + yb[m] + j) & 255].astype(np.int64)
    h = pm[r, (h + zb[m] + k) & 255].astype(np.int64)
    h = pm[r, (h + wb[m] + l) & 255].astype(np.int64)
    g = GRADIENTS_4D[h >> 2]
    a = attn[m]
    a = a * a
    value[m] += a * a * (g[:, 0] * dx[m] + g[:, 1] * dy[m] + g[:, 2] * dz[m] + g[:, 3] * dw[m])
  return (value / NORM_4D).reshape(shape)


class OpenSimplex:
  """Drop-in for `opensimplex.OpenSimplex` (only what the reference calls)."""

  def __init__(self, seed: int = 0):
    self._perm = make_perm(seed)

  def noise4d(self, x, y, z, w) -> float:
    return noise4d_scalar(self._perm, float(x), float(y), float(z), float(w))
