import numpy as _np
import jax.random as _jr
class distributions:
    class LogitNormal:
        def __init__(self, loc, scale): self.loc, self.scale = loc, scale
        def sample(self, seed):
            z = self.loc + self.scale * float(_jr.normal(seed))
            return _np.asarray(1.0 / (1.0 + _np.exp(-z)))
class bijectors: pass
