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
  * each SELECTED lattice vertex v contributes  max(0, 2 - |d|^2)^4 * (g(v) . d),
    d = displacement from the un-stretched vertex, g(v) one of 64 gradients
    (permutations of (+-3, +-1, +-1, +-1)) selected by
    perm[(perm[(perm[(perm[x&255]+y)&255]+z)&255]+w)&255] & 0xFC;
  * result = sum / 30.

VERTEX SELECTION (form='tree', the default since round 2).  The package does
not sum over every in-range vertex: `noise4d` walks a decision tree on the
position (xins, yins, zins, wins) inside the unit cell, inSum = their sum:

  region A  inSum <= 1      base = (0,0,0,0) + the 4 unit vertices
  region B  inSum >= 3      base = (1,1,1,1) + the 4 vertices with one 0
  region C  1 < inSum <= 2  base = the 4 unit + the 6 two-ones vertices
  region D  2 < inSum < 3   base = the 4 one-zero + the 6 two-ones vertices

plus THREE "extra" vertices chosen from the two closest base candidates
(`tree_extras_*` below spell the rule out per region; B and D are the
point-reflections v -> 1 - v of A and C, including the tie-breaking).  The
published source spells every case out coordinate by coordinate (~700 lines);
it is restated here by its vertex SETS: each case's three extras are written
as lattice offsets and every displacement follows from the one rule
d = d0 - v - SQUISH * sum(v), which is what the source's hand-expanded
dx_ext / dy_ext / ... constants evaluate to.  Two independent restatements are
kept and tested against each other: the scalar one writes all four regions out,
the vectorised one evaluates B and D through the reflection.

form='all' is the round-1 definition: the sum over EVERY vertex whose kernel is
non-zero (16 cube corners + 64 one-step-out candidates).  Measured over 8 M
uniform points (tests/test_oracle_golden.py): the two forms differ at 12 % of
the points, by at most 4.9e-4 (0.2 % of the noise's standard deviation, rms
8.7e-6) -- the tree leaves out only vertices whose kernel is almost zero -- and
their variances agree to 8 digits (0.0611).  The reference's constant
OPENSIMPLEX_VARIANCE = 0.0569 (env/simplex_wind_noise.py:69) is therefore NOT
explained by vertex selection (the base vertices alone give 0.0615): it is an
empirical constant of the reference, not the variance of noise4d over uniform
points, and cannot serve as a gate on the restatement.  The CUDA kernels
implement form='tree'; reference <-> oracle parity for noise VALUES stays
unpinned (no golden vector of the package exists in the reference).
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


# ---------------------------------------------------------------------------------------------
# Vertex selection of the published noise4d (see the module docstring).  Points are bit masks over
# the axes (bit 0 = x ... bit 3 = w); _vec(mask) is the lattice offset with a 1 on every set axis.

_UNIT = ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1))
_PAIR = ((1, 1, 0, 0), (1, 0, 1, 0), (1, 0, 0, 1), (0, 1, 1, 0), (0, 1, 0, 1), (0, 0, 1, 1))
_TRIPLE = ((1, 1, 1, 0), (1, 1, 0, 1), (1, 0, 1, 1), (0, 1, 1, 1))
# base contributions per region, in the order the published source adds them
BASE = {'A': ((0, 0, 0, 0),) + _UNIT, 'B': _TRIPLE + ((1, 1, 1, 1),), 'C': _UNIT + _PAIR, 'D': _TRIPLE + _PAIR}


def _vec(mask):
  return [(mask >> c) & 1 for c in range(4)]


def _axes(mask, want):
  """Axes (ascending) whose bit in `mask` equals `want`."""
  return [c for c in range(4) if ((mask >> c) & 1) == want]


def _bump(v, axis, by):
  v = list(v)
  v[axis] += by
  return tuple(v)


def tree_extras_scalar(ins):
  """(region, [3 extra lattice offsets]) for a point `ins` = (xins, yins, zins, wins) of the unit cell."""
  xins, yins, zins, wins = ins
  in_sum = xins + yins + zins + wins
  if in_sum <= 1:
    # inside the pentachoron at (0,0,0,0): the two closest of the unit vertices (largest coordinate = closest)
    a_po, a_sc, b_po, b_sc = 0x1, xins, 0x2, yins
    for sc, po in ((zins, 0x4), (wins, 0x8)):
      if a_sc >= b_sc and sc > b_sc:
        b_sc, b_po = sc, po
      elif a_sc < b_sc and sc > a_sc:
        a_sc, a_po = sc, po
    uins = 1 - in_sum
    if uins > a_sc or uins > b_sc:
      # (0,0,0,0) is one of the two closest: the other closest vertex c, with each of its zeros lowered to -1
      c = b_po if b_sc > a_sc else a_po
      ext = [_bump(_vec(c), ax, -1) for ax in _axes(c, 0)]
    else:
      # c = the vertex with both closest axes set; itself and its two zeros lowered to -1
      c = a_po | b_po
      ext = [tuple(_vec(c))] + [_bump(_vec(c), ax, -1) for ax in _axes(c, 0)]
    return 'A', ext
  if in_sum >= 3:
    # inside the pentachoron at (1,1,1,1): the two closest of the one-zero vertices (smallest coordinate = closest)
    a_po, a_sc, b_po, b_sc = 0xE, xins, 0xD, yins
    for sc, po in ((zins, 0xB), (wins, 0x7)):
      if a_sc <= b_sc and sc < b_sc:
        b_sc, b_po = sc, po
      elif a_sc > b_sc and sc < a_sc:
        a_sc, a_po = sc, po
    uins = 4 - in_sum
    if uins < a_sc or uins < b_sc:
      c = b_po if b_sc < a_sc else a_po
      ext = [_bump(_vec(c), ax, +1) for ax in _axes(c, 1)]
    else:
      c = a_po & b_po
      ext = [tuple(_vec(c))] + [_bump(_vec(c), ax, +1) for ax in _axes(c, 1)]
    return 'B', ext
  if in_sum <= 2:
    # first dispentachoron: the closer of each complementary two-ones pair, then the unit vertices
    a_sc, a_po = (xins + yins, 0x3) if xins + yins > zins + wins else (zins + wins, 0xC)
    b_sc, b_po = (xins + zins, 0x5) if xins + zins > yins + wins else (yins + wins, 0xA)
    a_big = b_big = True
    third = (xins + wins, 0x9) if xins + wins > yins + zins else (yins + zins, 0x6)
    cands = [(third[0], third[1], True)] + [(2 - in_sum + v, 1 << c, False) for c, v in enumerate(ins)]
    for sc, po, big in cands:
      if a_sc >= b_sc and sc > b_sc:
        b_sc, b_po, b_big = sc, po, big
      elif a_sc < b_sc and sc > a_sc:
        a_sc, a_po, a_big = sc, po, big
    if a_big and b_big:
      c1, c2 = a_po | b_po, a_po & b_po                 # three ones / the shared axis
      ext = [tuple(_vec(c1)), _bump(_vec(c1), _axes(c1, 0)[0], -1), _bump((0, 0, 0, 0), _axes(c2, 1)[0], 2)]
    elif not a_big and not b_big:
      c = a_po | b_po
      ext = [_bump(_vec(c), ax, -1) for ax in _axes(c, 0)] + [(0, 0, 0, 0)]
    else:
      c1, c2 = (a_po, b_po) if a_big else (b_po, a_po)  # the two-ones point / the unit point
      ext = [_bump(_vec(c1), ax, -1) for ax in _axes(c1, 0)] + [_bump((0, 0, 0, 0), _axes(c2, 1)[0], 2)]
    return 'C', ext
  # second dispentachoron
  a_sc, a_po = (xins + yins, 0xC) if xins + yins < zins + wins else (zins + wins, 0x3)
  b_sc, b_po = (xins + zins, 0xA) if xins + zins < yins + wins else (yins + wins, 0x5)
  a_big = b_big = True
  third = (xins + wins, 0x6) if xins + wins < yins + zins else (yins + zins, 0x9)
  cands = [(third[0], third[1], True)] + [(3 - in_sum + v, 0xF ^ (1 << c), False) for c, v in enumerate(ins)]
  for sc, po, big in cands:
    if a_sc <= b_sc and sc < b_sc:
      b_sc, b_po, b_big = sc, po, big
    elif a_sc > b_sc and sc < a_sc:
      a_sc, a_po, a_big = sc, po, big
  if a_big and b_big:
    c1, c2 = a_po & b_po, a_po | b_po                   # the shared axis / three ones
    ax = _axes(c1, 1)[0]
    ext = [_bump((0, 0, 0, 0), ax, 1), _bump((0, 0, 0, 0), ax, 2), _bump((1, 1, 1, 1), _axes(c2, 0)[0], -2)]
  elif not a_big and not b_big:
    c = a_po & b_po
    ext = [_bump(_vec(c), ax, +1) for ax in _axes(c, 1)] + [(1, 1, 1, 1)]
  else:
    c1, c2 = (a_po, b_po) if a_big else (b_po, a_po)    # the two-ones point / the one-zero point
    ext = [_bump(_vec(c1), ax, +1) for ax in _axes(c1, 1)] + [_bump((1, 1, 1, 1), _axes(c2, 0)[0], -2)]
  return 'D', ext


def noise4d_scalar(perm, x, y, z, w, form: str = 'tree') -> float:
  s = (x + y + z + w) * STRETCH_4D
  xs, ys, zs, ws = x + s, y + s, z + s, w + s
  xb, yb, zb, wb = (int(np.floor(v)) for v in (xs, ys, zs, ws))
  q = (xb + yb + zb + wb) * SQUISH_4D
  dx0, dy0, dz0, dw0 = x - (xb + q), y - (yb + q), z - (zb + q), w - (wb + q)
  if form == 'tree':
    region, ext = tree_extras_scalar((xs - xb, ys - yb, zs - zb, ws - wb))
    vertices = list(BASE[region]) + ext
  else:
    assert form == 'all', form
    vertices = CANDIDATES
  value = 0.0
  for (i, j, k, l) in vertices:
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


# ---- vectorised selection: regions B and D through the reflection v -> 1 - v of A and C -------------------------

_EYE = np.eye(4, dtype=np.int64)


def _bits(mask):
  return np.stack([(mask >> c) & 1 for c in range(4)], -1).astype(np.int64)


def _kth_axis(mask, want, k):
  """Index of the k-th (0-based) axis whose bit in `mask` equals `want` (vectorised)."""
  hit = _bits(mask) == want
  return np.argmax(hit & (np.cumsum(hit, -1) == k + 1), -1)


def _replace(state, score, point, big):
  a_sc, a_po, a_big, b_sc, b_po, b_big = state
  into_b = (a_sc >= b_sc) & (score > b_sc)
  into_a = (a_sc < b_sc) & (score > a_sc)
  return (np.where(into_a, score, a_sc), np.where(into_a, point, a_po), np.where(into_a, big, a_big),
          np.where(into_b, score, b_sc), np.where(into_b, point, b_po), np.where(into_b, big, b_big))


def _extras_low_pentachoron(u):
  x, y, z, w = u.T
  n = len(x)
  yes = np.ones(n, bool)
  st = (x, np.full(n, 1), yes, y, np.full(n, 2), yes)
  st = _replace(st, z, 4, yes)
  a_sc, a_po, _, b_sc, b_po, _ = _replace(st, w, 8, yes)
  origin_close = ((1 - (x + y + z + w)) > a_sc) | ((1 - (x + y + z + w)) > b_sc)
  ext = np.zeros((n, 3, 4), np.int64)
  c = np.where(b_sc > a_sc, b_po, a_po)
  for k in range(3):
    ext[origin_close, k] = (_bits(c) - _EYE[_kth_axis(c, 0, k)])[origin_close]
  c = a_po | b_po
  rest = ~origin_close
  ext[rest, 0] = _bits(c)[rest]
  ext[rest, 1] = (_bits(c) - _EYE[_kth_axis(c, 0, 0)])[rest]
  ext[rest, 2] = (_bits(c) - _EYE[_kth_axis(c, 0, 1)])[rest]
  return ext


def _extras_low_dispentachoron(u):
  x, y, z, w = u.T
  n = len(x)
  yes, no = np.ones(n, bool), np.zeros(n, bool)
  total = x + y + z + w
  c = x + y > z + w
  st = (np.where(c, x + y, z + w), np.where(c, 0x3, 0xC), yes)
  c = x + z > y + w
  st = st + (np.where(c, x + z, y + w), np.where(c, 0x5, 0xA), yes)
  c = x + w > y + z
  st = _replace(st, np.where(c, x + w, y + z), np.where(c, 0x9, 0x6), yes)
  for axis, v in enumerate((x, y, z, w)):
    st = _replace(st, 2 - total + v, 1 << axis, no)
  _, a_po, a_big, _, b_po, b_big = st
  ext = np.zeros((n, 3, 4), np.int64)
  m = a_big & b_big
  c1, c2 = a_po | b_po, a_po & b_po
  ext[m, 0] = _bits(c1)[m]
  ext[m, 1] = (_bits(c1) - _EYE[_kth_axis(c1, 0, 0)])[m]
  ext[m, 2] = (2 * _EYE[_kth_axis(c2, 1, 0)])[m]
  m = ~a_big & ~b_big
  c = a_po | b_po
  ext[m, 0] = (_bits(c) - _EYE[_kth_axis(c, 0, 0)])[m]
  ext[m, 1] = (_bits(c) - _EYE[_kth_axis(c, 0, 1)])[m]
  ext[m, 2] = 0
  m = a_big ^ b_big
  c1, c2 = np.where(a_big, a_po, b_po), np.where(a_big, b_po, a_po)
  ext[m, 0] = (_bits(c1) - _EYE[_kth_axis(c1, 0, 0)])[m]
  ext[m, 1] = (_bits(c1) - _EYE[_kth_axis(c1, 0, 1)])[m]
  ext[m, 2] = (2 * _EYE[_kth_axis(c2, 1, 0)])[m]
  return ext


def tree_vertices(ins):
  """ins float[n, 4] in the unit cell -> (offsets int64[n, 13, 4], valid bool[n, 13]): base vertices of the region,
  then the three extras."""
  ins = np.asarray(ins, np.float64)
  n = len(ins)
  total = ins[:, 0] + ins[:, 1] + ins[:, 2] + ins[:, 3]
  region = np.where(total <= 1, 0, np.where(total >= 3, 1, np.where(total <= 2, 2, 3)))
  verts = np.zeros((n, 13, 4), np.int64)
  valid = np.zeros((n, 13), bool)
  for r, name in enumerate('ABCD'):
    m = np.nonzero(region == r)[0]
    if m.size == 0:
      continue
    reflect = name in 'BD'
    u = 1 - ins[m] if reflect else ins[m]
    ext = _extras_low_pentachoron(u) if name in 'AB' else _extras_low_dispentachoron(u)
    if reflect:
      ext = 1 - ext
    base = np.array(BASE[name], np.int64)
    nb = len(base)
    verts[m, :nb] = base
    verts[m[:, None], np.arange(nb, nb + 3)[None, :]] = ext
    valid[m, :nb + 3] = True
  return verts, valid


def noise4d(perm, x, y, z, w, form: str = 'tree') -> np.ndarray:
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
  if form == 'tree':
    verts, valid = tree_vertices(np.stack([xs - xb, ys - yb, zs - zb, ws - wb], -1))
    slots = [(verts[:, k, 0], verts[:, k, 1], verts[:, k, 2], verts[:, k, 3], valid[:, k]) for k in range(13)]
  else:
    assert form == 'all', form
    slots = [(i, j, k, l, True) for (i, j, k, l) in CANDIDATES]
  for (i, j, k, l, ok) in slots:
    t = (i + j + k + l) * SQUISH_4D
    dx, dy, dz, dw = dx0 - i - t, dy0 - j - t, dz0 - k - t, dw0 - l - t
    attn = 2.0 - dx * dx - dy * dy - dz * dz - dw * dw
    m = (attn > 0.0) & ok
    if not m.any():
      continue
    r = rows[m]
    pick = (lambda o: o[m]) if form == 'tree' else (lambda o: o)
    h = pm[r, (xb[m] + pick(i)) & 255].astype(np.int64)
    h = pm[r, (h + yb[m] + pick(j)) & 255].astype(np.int64)
    h = pm[r, (h + zb[m] + pick(k)) & 255].astype(np.int64)
    h = pm[r, (h + wb[m] + pick(l)) & 255].astype(np.int64)
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
