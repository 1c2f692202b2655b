"""Tier-0 harness bootstrap: makes the UNMODIFIED reference importable in this container.

Usage (this container only; /root/reference does not exist on the GPU box):

    import tests.golden.tier0.boot  # noqa  (must come before any reference import)

What it does:
  * puts `shim/` (stand-ins for jax, flax, gin, gym, s2sphere, transitions,
    tensorflow, tensorflow_probability, opensimplex -- all missing here) and
    /root/reference on sys.path;
  * restores `scipy.interpolate.interp2d` (removed in SciPy >= 1.14, used at import
    time by env/balloon/acs.py:31-42) as the mathematically identical bilinear
    interpolant with nearest/clamp extrapolation;
  * gives the `units` value classes a `__hash__` so Python >= 3.11 dataclasses accept
    them as defaults (env/balloon/balloon.py:167-198).
No reference file is edited or copied.
"""
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.abspath(os.path.join(_HERE, '..', '..', '..'))
REFERENCE_ROOT = os.environ.get('BLE_REFERENCE_ROOT', '/root/reference')

if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'balloon_learning_environment')):
  raise ImportError(f'Tier-0 harness needs the reference checkout at {REFERENCE_ROOT}')

for p in (REFERENCE_ROOT, os.path.join(_HERE, 'shim'), _REPO):
  if p in sys.path:
    sys.path.remove(p)
  sys.path.insert(0, p)
sys.dont_write_bytecode = True
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import scipy.interpolate as _si  # noqa: E402
from scipy.interpolate import RectBivariateSpline  # noqa: E402


class _Interp2d:
  """Bilinear interp2d(x, y, z) with clamp outside the grid (FITPACK kx=ky=1 behaviour)."""

  def __init__(self, x, y, z, fill_value=None, kind='linear'):
    self.x, self.y = np.asarray(x), np.asarray(y)
    self.s = RectBivariateSpline(
        self.x, self.y, np.asarray(z).reshape(len(y), len(x)).T, kx=1, ky=1)

  def __call__(self, x, y):
    x = min(max(float(x), self.x[0]), self.x[-1])
    y = min(max(float(y), self.y[0]), self.y[-1])
    return np.array([self.s(x, y)[0, 0]])


_si.interp2d = _Interp2d

from balloon_learning_environment.utils import units as _u  # noqa: E402

for _c in (_u.Power, _u.Energy, _u.Distance, _u.Velocity):
  _c.__hash__ = lambda self: id(self)
