import numpy as _np
from numpy import *  # noqa
ndarray = _np.ndarray
pi = _np.pi
int32 = _np.int32
float32 = _np.float32
def linspace(start, stop, num, dtype=None):
    out = _np.linspace(start, stop, num)
    return out.astype(dtype if dtype is not None else _np.float32)
