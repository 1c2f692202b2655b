# PROBE-ONLY stub of jax (numpy-backed).
import numpy as _np
from . import numpy  # noqa
from . import random  # noqa
def jit(f=None, **kw):
    if f is None: return lambda g: g
    return f
def tree_map(f, t): raise NotImplementedError
def _resize_axis_linear(x, out, axis):
    n = x.shape[axis]; scale = n / out
    pos = (_np.arange(out) + 0.5) * scale - 0.5          # half-pixel centres
    i0 = _np.floor(pos).astype(int); w1 = (pos - i0).astype(_np.float32); w0 = 1 - w1
    a = _np.clip(i0, 0, n - 1); b = _np.clip(i0 + 1, 0, n - 1)  # == drop OOB taps + renormalise for a triangle kernel
    xa = _np.take(x, a, axis=axis); xb = _np.take(x, b, axis=axis)
    sh = [1] * x.ndim; sh[axis] = out
    return xa * w0.reshape(sh) + xb * w1.reshape(sh)
class image:
    @staticmethod
    def resize(x, shape, method='linear'):
        assert method == 'linear'
        y = _np.asarray(x, _np.float32)
        for ax, (o, n) in enumerate(zip(shape, y.shape)):
            if o != n: y = _resize_axis_linear(y, o, ax)
        return y.astype(_np.float32)
