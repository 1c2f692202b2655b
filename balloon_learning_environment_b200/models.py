"""Wind-field decoder weights (models/models.py:33-77: load_offlineskies22).

The reference ships `models/offlineskies22_decoder.msgpack` (a flax msgpack tree); this loader reads that
file format with `msgpack` alone.  Without a path it returns random-init weights of the same architecture
(generative/vae.py:134-186) so that benchmarks and tests have a decoder to run.
"""
from typing import Dict, Optional

import numpy as np

LAYER_SIZES = (64, 1000, 1000, 1000, 4410)


def load_decoder(path: Optional[str] = None, seed: int = 22) -> Dict[str, Dict[str, np.ndarray]]:
  if path:
    import msgpack

    def ext_hook(code, data):
      if code == 1:                                    # flax ndarray extension: (shape, dtype name, bytes)
        shape, dtype, buf = msgpack.unpackb(data, raw=False)
        return np.frombuffer(buf, dtype=np.dtype(dtype)).reshape(shape)
      return msgpack.ExtType(code, data)

    with open(path, 'rb') as f:
      tree = msgpack.unpackb(f.read(), ext_hook=ext_hook, raw=False)
    return tree['params'] if 'params' in tree else tree
  rng = np.random.default_rng(seed)
  return {f'Dense_{i}': {'kernel': (rng.standard_normal((LAYER_SIZES[i], LAYER_SIZES[i + 1]))
                                    * (2.0 / LAYER_SIZES[i]) ** 0.5).astype(np.float32),
                         'bias': np.zeros(LAYER_SIZES[i + 1], np.float32)} for i in range(4)}
