"""Wind-field VAE decoder (ORACLE / test infrastructure).  Follows generative/vae.py:134-186 and
env/generative_wind_field.py:52-62.

64 latents -> 3 x (Dense 1000 + ReLU) -> Dense 4410 -> reshape (7, 7, 90) -> `jax.image.resize`
('linear') to (23, 23, 90) -> central differences (stream function -> incompressible wind) ->
crop to (21, 21, 90) -> u = dPsi/d(axis 0), v = -dPsi/d(axis 1) -> [21, 21, 10, 9, 2].

NUMERICAL PARITY OF THE RESIZE IS UNPINNED: jax is not installed here and the reference only
tests the output shape (generative/vae_test.py:43-54).  `resize_linear` restates jax.image.resize's
documented behaviour for up-sampling (half-pixel centres, triangle kernel, out-of-range taps dropped
and the weights renormalised == edge clamp).  fp32 throughout, as JAX computes by default.
"""
import numpy as np

HIDDEN = 1000
LATENTS = 64
FLOW_W = 7
FLOW_FIELDS = 90            # pressure_slices * time_slices = 10 * 9
OUT_UNITS = FLOW_W * FLOW_W * FLOW_FIELDS     # 4410


def resize_weights(n_in: int, n_out: int) -> np.ndarray:
  """[n_out, n_in] float32 interpolation matrix of jax.image.resize(method='linear'), up-sampling."""
  scale = n_out / n_in
  pos = (np.arange(n_out) + 0.5) / scale - 0.5
  w = np.maximum(0.0, 1.0 - np.abs(pos[:, None] - np.arange(n_in)[None, :]))
  w = w / w.sum(axis=1, keepdims=True)
  return w.astype(np.float32)


def decode(params, z: np.ndarray) -> np.ndarray:
  """params: dict Dense_0..Dense_3 -> {kernel [in, out], bias [out]} (flax layout); z [F, 64] ->
  float32 [F, 21, 21, 10, 9, 2]."""
  x = np.asarray(z, np.float32)
  for i in range(3):                                                       # vae.py:142-144
    p = params[f'Dense_{i}']
    x = np.maximum(x @ np.asarray(p['kernel'], np.float32) + np.asarray(p['bias'], np.float32), 0.0)
  p = params['Dense_3']
  x = x @ np.asarray(p['kernel'], np.float32) + np.asarray(p['bias'], np.float32)   # :145
  f = x.shape[0]
  flow = x.reshape(f, FLOW_W, FLOW_W, FLOW_FIELDS)                          # :149-151
  w = resize_weights(FLOW_W, 23)
  flow = np.einsum('ai,fijc->fajc', w, flow).astype(np.float32)             # :160-164 (axis 0)
  flow = np.einsum('bj,fajc->fabc', w, flow).astype(np.float32)             #          (axis 1)
  dy = (np.roll(flow, -1, axis=1) - np.roll(flow, 1, axis=1)) / np.float32(2.0)   # :171-173
  dx = (np.roll(flow, -1, axis=2) - np.roll(flow, 1, axis=2)) / np.float32(2.0)   # :174-176
  dy = dy[:, 1:-1, 1:-1, :]
  dx = dx[:, 1:-1, 1:-1, :]
  u = dy.reshape(f, 21, 21, 10, 9)                                          # :182
  v = -dx.reshape(f, 21, 21, 10, 9)                                         # :183
  return np.stack([u, v], axis=-1).astype(np.float32)                       # :186


def synthetic_params(seed: int = 0):
  """Random-init decoder weights of the reference architecture (He-style scale so that decoded
  winds come out at a few m/s); used where the 25.9 MB offlineskies22 checkpoint cannot travel."""
  rng = np.random.default_rng(seed)
  dims = [LATENTS, HIDDEN, HIDDEN, HIDDEN, OUT_UNITS]
  params = {}
  for i in range(4):
    scale = np.sqrt(2.0 / dims[i]) * (6.0 if i == 3 else 1.0)
    params[f'Dense_{i}'] = {
        'kernel': (rng.standard_normal((dims[i], dims[i + 1])) * scale).astype(np.float32),
        'bias': (rng.standard_normal(dims[i + 1]) * 0.05).astype(np.float32)}
  return params
