# PROBE-ONLY functional mini-linen: enough for vae.Decoder().apply(params, z).
import numpy as np
_CTX = []
class Module:
    def __init__(self, *a, **k):
        ann = {}
        for c in reversed(type(self).__mro__): ann.update(getattr(c, '__annotations__', {}))
        names = list(ann)
        for n, v in zip(names, a): setattr(self, n, v)
        for n, v in k.items(): setattr(self, n, v)
    def apply(self, params, *args, **kw):
        _CTX.append({'params': params['params'], 'n': 0})
        try: return self(*args, **kw)
        finally: _CTX.pop()
def compact(f): return f
def relu(x): return np.maximum(x, 0)
class Dense:
    def __init__(self, features, name=None, **k): self.features, self.name = features, name
    def __call__(self, x):
        c = _CTX[-1]; name = self.name or f"Dense_{c['n']}"
        if self.name is None: c['n'] += 1
        p = c['params'][name]
        return (np.asarray(x, np.float32) @ p['kernel'] + p['bias']).astype(np.float32)
