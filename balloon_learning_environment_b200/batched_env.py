"""Batched, GPU-resident Balloon Learning Environment (host side).

`BatchedBalloonArena` mirrors the reference's `BalloonArena` (env/balloon_arena.py:123-275) and
`BatchedBalloonEnv` its `BalloonEnv` (env/balloon_env.py:105-300) for N balloons at once.  All
numerics run in hand-written CUDA behind the C ABI of include/ble_b200.h; PyTorch only owns the
device buffers and the stream.  There is no CPU path.
"""
import ctypes
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from balloon_learning_environment_b200 import _lib
from balloon_learning_environment_b200 import sharding

# AltitudeControlCommand (env/balloon/control.py:21-25) and BalloonStatus (env/balloon/balloon.py:66-70)
DOWN, STAY, UP = 0, 1, 2
STATUS_OK, STATUS_OUT_OF_POWER, STATUS_BURST, STATUS_ZEROPRESSURE = 0, 1, 2, 3
FIELD_SHAPE = (21, 21, 10, 9, 2)     # generative/vae.py:38-51


class Discrete:
  """Minimal stand-in for gym.spaces.Discrete (env/balloon_env.py:262-264)."""

  def __init__(self, n: int):
    self.n = n

  def __repr__(self):
    return f'Discrete({self.n})'


def _ptr(t: Optional[torch.Tensor]):
  return ctypes.c_void_p(0 if t is None else t.data_ptr())


class BatchedBalloonArena:
  """N balloons, each flying in its own (or a shared) wind field.  Device-resident state.

  Reference: BalloonArena (env/balloon_arena.py:123-275).  Differences that come with batching:
  `step` takes a tensor of N commands; a balloon whose status is not OK is left untouched
  (the reference asserts, env/balloon/balloon.py:288).
  """

  def __init__(self, num_envs: int, *, device: str = 'cuda:0', precision: str = 'fp32',
               wind_model: str = 'grid', enable_noise: bool = True, field_layout: str = 'x64',
               enable_features: bool = False, decoder_precision: str = 'tf32', auto_reset: bool = False):
    if not torch.cuda.is_available():
      raise _lib.BleError('BatchedBalloonArena needs a CUDA device (no CPU fallback exists)')
    self._lib = _lib.load()
    self.device = torch.device(device)
    self.num_envs = int(num_envs)
    self.precision = precision
    self.wind_model = wind_model
    self.enable_noise = bool(enable_noise)
    cfg = _lib.BleConfig(_lib.PRECISION[precision], _lib.WIND_MODEL[wind_model], int(enable_noise),
                         _lib.FIELD_LAYOUT[field_layout], int(enable_features), {'tf32': 0, 'fp32': 1}[decoder_precision],
                         int(bool(auto_reset)))
    self.auto_reset = bool(auto_reset)
    self.enable_features = bool(enable_features)
    handle = ctypes.c_void_p()
    dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
    rc = self._lib.ble_create(dev_index, self.num_envs, ctypes.byref(cfg), ctypes.byref(handle))
    _lib.check(self._lib, None, rc, 'ble_create')
    self._h = handle
    n = self.num_envs
    with torch.cuda.device(self.device):
      self._reward = torch.zeros(n, dtype=torch.float32, device=self.device)
      self._done = torch.zeros(n, dtype=torch.uint8, device=self.device)
      self._wind = torch.zeros(n, 2, dtype=torch.float32, device=self.device)
      self._status = torch.zeros(n, dtype=torch.uint8, device=self.device)
      self._time_elapsed = torch.zeros(n, dtype=torch.int32, device=self.device)
      self._sim_error = torch.zeros(n, dtype=torch.uint8, device=self.device)
    self._step_out = _lib.BleStepOut(self._reward.data_ptr(), self._done.data_ptr(), self._wind.data_ptr(),
                                     self._status.data_ptr(), self._time_elapsed.data_ptr(), self._sim_error.data_ptr())
    self._keepalive = []

  # -- plumbing ---------------------------------------------------------------------------------
  def _stream(self):
    return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  def _check(self, rc, what):
    _lib.check(self._lib, self._h, rc, what)

  def close(self):
    if getattr(self, '_h', None) is not None:
      torch.cuda.synchronize(self.device)
      self._lib.ble_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass

  @property
  def launch_count(self) -> int:
    return int(self._lib.ble_launch_count(self._h))

  # -- wind -------------------------------------------------------------------------------------
  def set_wind_fields(self, fields: torch.Tensor, env_to_field: Optional[torch.Tensor] = None):
    """fields: float32 [F,21,21,10,9,2] (GridWindFieldSampler.sample_field layout); env_to_field int32 [N]."""
    fields = fields.to(self.device, torch.float32).contiguous()
    if tuple(fields.shape[1:]) != FIELD_SHAPE:
      raise ValueError(f'fields must have shape [F, {FIELD_SHAPE}], got {tuple(fields.shape)}')
    if env_to_field is not None:
      env_to_field = env_to_field.to(self.device, torch.int32).contiguous()
      if env_to_field.numel() != self.num_envs:
        raise ValueError('env_to_field must have one entry per balloon')
      if int(env_to_field.max()) >= fields.shape[0] or int(env_to_field.min()) < 0:
        raise ValueError('env_to_field refers to a field that was not provided')
    self._n_fields = int(fields.shape[0])
    rc = self._lib.ble_upload_fields(self._h, _ptr(fields), fields.shape[0], _ptr(env_to_field), self._stream())
    self._check(rc, 'ble_upload_fields')
    self._keepalive = [fields, env_to_field]

  def alloc_wind_fields(self, n_fields: int):
    """Room for n_fields grids; fill with write_wind_fields (keeps peak memory at one chunk)."""
    self._check(self._lib.ble_alloc_fields(self._h, int(n_fields), self._stream()), 'ble_alloc_fields')
    self._n_fields = int(n_fields)

  def write_wind_fields(self, fields: torch.Tensor, first_field: int):
    fields = fields.to(self.device, torch.float32).contiguous()
    if tuple(fields.shape[1:]) != FIELD_SHAPE:
      raise ValueError(f'fields must have shape [F, {FIELD_SHAPE}], got {tuple(fields.shape)}')
    rc = self._lib.ble_write_fields(self._h, _ptr(fields), int(first_field), fields.shape[0], self._stream())
    self._check(rc, 'ble_write_fields')
    torch.cuda.current_stream(self.device).synchronize()

  def set_field_map(self, env_to_field: torch.Tensor):
    env_to_field = env_to_field.to(self.device, torch.int32).contiguous()
    if env_to_field.numel() != self.num_envs:
      raise ValueError('env_to_field must have one entry per balloon')
    if int(env_to_field.max()) >= self._n_fields or int(env_to_field.min()) < 0:
      raise ValueError('env_to_field refers to a field that was not allocated')
    self._check(self._lib.ble_set_field_map(self._h, _ptr(env_to_field), self._stream()), 'ble_set_field_map')
    torch.cuda.current_stream(self.device).synchronize()
    self._keepalive = [None, env_to_field]

  # -- VAE wind generator (reset path) ------------------------------------------------------------
  def set_decoder(self, params) -> None:
    """params: {'Dense_0'..'Dense_3': {'kernel': [in, out], 'bias': [out]}} (the flax tree of
    models/offlineskies22_decoder.msgpack['params'], numpy or torch)."""
    ks, bs = [], []
    for i in range(4):
      layer = params[f'Dense_{i}']
      as_dev = lambda a: (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, np.float32))).to(self.device, torch.float32).contiguous()
      ks.append(as_dev(layer['kernel']))
      bs.append(as_dev(layer['bias']))
    expected = [(64, 1000), (1000, 1000), (1000, 1000), (1000, 4410)]
    if [tuple(k.shape) for k in ks] != expected or [b.numel() for b in bs] != [1000, 1000, 1000, 4410]:
      raise ValueError('decoder weights do not have the vae.Decoder shapes')
    kp = (ctypes.c_void_p * 4)(*[k.data_ptr() for k in ks])
    bp = (ctypes.c_void_p * 4)(*[b.data_ptr() for b in bs])
    self._check(self._lib.ble_set_decoder(self._h, kp, bp, self._stream()), 'ble_set_decoder')
    torch.cuda.current_stream(self.device).synchronize()

  def decode_wind_fields(self, latents: torch.Tensor) -> torch.Tensor:
    """vae.Decoder().apply(params, z) for a batch: latents float32 [F, 64] -> [F,21,21,10,9,2]."""
    latents = latents.to(self.device, torch.float32).contiguous()
    if latents.dim() != 2 or latents.shape[1] != 64:
      raise ValueError('latents must be [F, 64]')
    out = torch.empty(latents.shape[0], *FIELD_SHAPE, dtype=torch.float32, device=self.device)
    rc = self._lib.ble_decode_fields(self._h, _ptr(latents), latents.shape[0], _ptr(out), self._stream())
    self._check(rc, 'ble_decode_fields')
    return out

  def generate_wind_fields(self, n_fields: int, seed: int, chunk: int = 2048) -> None:
    """GenerativeWindFieldSampler.sample_field for n_fields grids: z ~ N(0, I_64) -> decoder -> field bank."""
    g = torch.Generator(device=self.device)
    g.manual_seed(int(seed))
    self.alloc_wind_fields(n_fields)
    for first in range(0, n_fields, chunk):
      c = min(chunk, n_fields - first)
      z = torch.randn(c, 64, generator=g, device=self.device, dtype=torch.float32)
      self.write_wind_fields(self.decode_wind_fields(z), first)

  def sample_wind_fields(self, seeds: torch.Tensor, first_field: int = 0) -> None:
    """GenerativeWindFieldSampler.sample_field for one seed per field: Philox(seed) latents -> decoder ->
    fields [first_field, first_field + len(seeds)) of the bank (call alloc_wind_fields first)."""
    seeds = seeds.to(self.device, torch.int64).contiguous()
    rc = self._lib.ble_generate_fields(self._h, _ptr(seeds), int(first_field), seeds.numel(), self._stream())
    self._check(rc, 'ble_generate_fields')

  def sample_latents(self, seeds: torch.Tensor) -> torch.Tensor:
    """The latents sample_wind_fields decodes for these seeds: int64 [K] -> float32 [K, 64], z ~ N(0, I)."""
    seeds = seeds.to(self.device, torch.int64).contiguous()
    out = torch.empty(seeds.numel(), 64, dtype=torch.float32, device=self.device)
    self._check(self._lib.ble_sample_latents(self._h, _ptr(seeds), seeds.numel(), _ptr(out), self._stream()), 'ble_sample_latents')
    return out

  def sample_wind_fields_at(self, seeds: torch.Tensor, field_index: torch.Tensor) -> None:
    """sample_wind_fields for a scattered set of fields: seeds int64 [K], field_index int32 [K] (distinct)."""
    seeds = seeds.to(self.device, torch.int64).contiguous()
    field_index = field_index.to(self.device, torch.int32).contiguous()
    if seeds.numel() != field_index.numel():
      raise ValueError('one seed per field index')
    if seeds.numel() == 0:
      return
    rc = self._lib.ble_generate_fields_at(self._h, _ptr(seeds), _ptr(field_index), seeds.numel(), self._stream())
    self._check(rc, 'ble_generate_fields_at')

  # -- evaluation surface (eval/eval_lib.py, agents/) ----------------------------------------------
  def station_seeker_actions(self, obs: torch.Tensor, with_level: bool = False):
    """StationSeekerAgent.pick_action for all balloons: obs float32 [N,1099] -> int32 [N]."""
    obs = self._check_obs(obs)
    actions = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
    best = torch.empty(self.num_envs, dtype=torch.int32, device=self.device) if with_level else None
    rc = self._lib.ble_agent_station_seeker(self._h, _ptr(obs), _ptr(actions), _ptr(best), self._stream())
    self._check(rc, 'ble_agent_station_seeker')
    return (actions, best) if with_level else actions

  def random_walk_actions(self, obs: torch.Tensor, seeds: torch.Tensor, step_index: int) -> torch.Tensor:
    """RandomWalkAgent: step_index 0 = begin_episode, k = k-th step; seeds int64 [N]."""
    obs = self._check_obs(obs)
    seeds = seeds.to(self.device, torch.int64).contiguous()
    if seeds.numel() != self.num_envs:
      raise ValueError('seeds must have one entry per balloon')
    actions = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
    rc = self._lib.ble_agent_random_walk(self._h, _ptr(obs), _ptr(seeds), int(step_index), _ptr(actions), self._stream())
    self._check(rc, 'ble_agent_random_walk')
    return actions

  def _check_obs(self, obs: torch.Tensor) -> torch.Tensor:
    if obs.shape != (self.num_envs, 1099):
      raise ValueError(f'obs must be [{self.num_envs}, 1099]')
    return obs.to(self.device, torch.float32).contiguous()

  def eval_begin(self) -> None:
    self._check(self._lib.ble_eval_begin(self._h, self._stream()), 'ble_eval_begin')

  def eval_accumulate(self, reward: torch.Tensor, flight_path: Optional[torch.Tensor] = None) -> None:
    """One pass of eval_agent's loop body after a step; flight_path: float32 [6, N] slot or None."""
    if flight_path is not None and (flight_path.shape != (6, self.num_envs) or flight_path.dtype != torch.float32
                                    or not flight_path.is_contiguous()):
      raise ValueError('flight_path must be a contiguous float32 [6, N] tensor')
    rc = self._lib.ble_eval_accumulate(self._h, _ptr(reward), _ptr(flight_path), self._stream())
    self._check(rc, 'ble_eval_accumulate')

  def eval_results(self) -> Dict[str, torch.Tensor]:
    out = torch.empty(len(_lib.E_ROWS), self.num_envs, dtype=torch.float64, device=self.device)
    self._check(self._lib.ble_eval_results(self._h, _ptr(out), self._stream()), 'ble_eval_results')
    return {k: out[i] for i, k in enumerate(_lib.E_ROWS)}

  def set_wind_noise(self, seeds: torch.Tensor, offsets: torch.Tensor):
    """seeds int64 [N,2,5], offsets float32 [N,2,5,4] (env/simplex_wind_noise.py:98-114)."""
    seeds = seeds.to(self.device, torch.int64).contiguous()
    offsets = offsets.to(self.device, torch.float32).contiguous()
    if tuple(seeds.shape) != (self.num_envs, 2, 5) or tuple(offsets.shape) != (self.num_envs, 2, 5, 4):
      raise ValueError('seeds must be [N,2,5] and offsets [N,2,5,4]')
    self._check(self._lib.ble_set_noise(self._h, _ptr(seeds), _ptr(offsets), self._stream()), 'ble_set_noise')
    torch.cuda.current_stream(self.device).synchronize()

  def wind_forecast(self, xyzt: torch.Tensor, field_idx: torch.Tensor) -> torch.Tensor:
    """GridBasedWindField.get_forecast for M points: xyzt float32 [M,4] (x km, y km, Pa, hours)."""
    xyzt = xyzt.to(self.device, torch.float32).contiguous()
    field_idx = field_idx.to(self.device, torch.int32).contiguous()
    m = xyzt.shape[0]
    uv = torch.empty(m, 2, dtype=torch.float32, device=self.device)
    rc = self._lib.ble_wind_gather(self._h, _ptr(xyzt), _ptr(field_idx), _ptr(uv), m, self._stream())
    self._check(rc, 'ble_wind_gather')
    return uv

  def wind_query(self, xyzt: torch.Tensor, env_idx: torch.Tensor, with_noise: bool) -> torch.Tensor:
    """WindField.get_forecast (with_noise=False) / get_ground_truth (True) at M points of given balloons' own
    fields and noise generators: xyzt float64 [M,4] (x m, y m, Pa, elapsed s), env_idx int32 [M] -> float32 [M,2]."""
    xyzt = xyzt.to(self.device, torch.float64).contiguous()
    env_idx = env_idx.to(self.device, torch.int32).contiguous()
    m = xyzt.shape[0]
    uv = torch.empty(m, 2, dtype=torch.float32, device=self.device)
    rc = self._lib.ble_wind_query(self._h, _ptr(xyzt), _ptr(env_idx), int(bool(with_noise)), _ptr(uv), m, self._stream())
    self._check(rc, 'ble_wind_query')
    return uv

  def atmosphere_query(self, which: str, q: torch.Tensor, env_idx: torch.Tensor) -> torch.Tensor:
    """Atmosphere.at_pressure / at_height for given balloons' atmospheres: q float64 [M] -> float64 [M,4]
    (height m, temperature K, pressure Pa, density kg/m^3); NaN rows where the reference asserts."""
    q = q.to(self.device, torch.float64).contiguous()
    env_idx = env_idx.to(self.device, torch.int32).contiguous()
    out = torch.empty(q.shape[0], 4, dtype=torch.float64, device=self.device)
    rc = self._lib.ble_atmosphere_query(self._h, {'pressure': 0, 'height': 1}[which], _ptr(q), _ptr(env_idx), _ptr(out),
                                        q.shape[0], self._stream())
    self._check(rc, 'ble_atmosphere_query')
    return out

  def wind_at_balloon(self) -> torch.Tensor:
    """Ground-truth wind at the balloons' current state, float32 [N,2] (get_measurements)."""
    uv = torch.empty(self.num_envs, 2, dtype=torch.float32, device=self.device)
    self._check(self._lib.ble_wind_at_balloon(self._h, _ptr(uv), self._stream()), 'ble_wind_at_balloon')
    return uv

  # -- state ------------------------------------------------------------------------------------
  def set_state(self, f64: torch.Tensor, i64: torch.Tensor):
    """set_balloon_state for all balloons: f64 [18,N] / i64 [10,N] in the row order of _lib.F_ROWS/I_ROWS."""
    f64 = f64.to(self.device, torch.float64).contiguous()
    i64 = i64.to(self.device, torch.int64).contiguous()
    if tuple(f64.shape) != (len(_lib.F_ROWS), self.num_envs) or tuple(i64.shape) != (len(_lib.I_ROWS), self.num_envs):
      raise ValueError('state matrices have the wrong shape')
    soa = _lib.BleStateSoa(f64.data_ptr(), i64.data_ptr())
    self._check(self._lib.ble_state_upload(self._h, ctypes.byref(soa), self._stream()), 'ble_state_upload')
    torch.cuda.current_stream(self.device).synchronize()

  def get_state(self) -> Tuple[torch.Tensor, torch.Tensor]:
    f64 = torch.empty(len(_lib.F_ROWS), self.num_envs, dtype=torch.float64, device=self.device)
    i64 = torch.empty(len(_lib.I_ROWS), self.num_envs, dtype=torch.int64, device=self.device)
    soa = _lib.BleStateSoa(f64.data_ptr(), i64.data_ptr())
    self._check(self._lib.ble_state_download(self._h, ctypes.byref(soa), self._stream()), 'ble_state_download')
    return f64, i64

  def get_state_dict(self) -> Dict[str, torch.Tensor]:
    f64, i64 = self.get_state()
    out = {k: f64[r] for r, k in enumerate(_lib.F_ROWS)}
    out.update({k: i64[r] for r, k in enumerate(_lib.I_ROWS)})
    return out

  def get_derived(self) -> Dict[str, torch.Tensor]:
    """BalloonState's derived properties (latlng, excess_energy, navigation_is_paused, ...)."""
    out = torch.empty(len(_lib.D_ROWS), self.num_envs, dtype=torch.float64, device=self.device)
    self._check(self._lib.ble_derived(self._h, _ptr(out), self._stream()), 'ble_derived')
    return {k: out[r] for r, k in enumerate(_lib.D_ROWS)}

  # -- observation surface ----------------------------------------------------------------------
  def features(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PerciatelliFeatureConstructor.get_features() for every balloon: float32 [N, 1099]."""
    if out is None:
      out = torch.empty(self.num_envs, 1099, dtype=torch.float32, device=self.device)
    self._check(self._lib.ble_features_perciatelli(self._h, _ptr(out), self._stream()), 'ble_features_perciatelli')
    return out

  def features_observe(self):
    self._check(self._lib.ble_features_observe(self._h, self._stream()), 'ble_features_observe')

  def features_track(self, on: bool) -> None:
    """on=False: reset / step stop appending WindGP measurements (rollouts that never read the observation)."""
    self._check(self._lib.ble_features_track(self._h, int(bool(on))), 'ble_features_track')

  def features_clear(self):
    self._check(self._lib.ble_features_clear(self._h, self._stream()), 'ble_features_clear')

  def init_derived(self, run_stable_init: bool = True):
    """Recompute power-safety sunrise/sunset (+ stable init) from the uploaded state."""
    self._check(self._lib.ble_init_derived(self._h, int(run_stable_init), self._stream()), 'ble_init_derived')

  # -- reset / step -----------------------------------------------------------------------------
  def reset(self, seeds: torch.Tensor, mask: Optional[torch.Tensor] = None):
    """BalloonArena.reset for every balloon (or those with mask != 0), one 64-bit seed each."""
    seeds = seeds.to(self.device).contiguous()
    if seeds.dtype not in (torch.int64, torch.uint64) or seeds.numel() != self.num_envs:
      raise ValueError('seeds must be an int64 tensor with one entry per balloon')
    if mask is not None:
      mask = mask.to(self.device, torch.uint8).contiguous()
    self._check(self._lib.ble_reset(self._h, _ptr(seeds), _ptr(mask), self._stream()), 'ble_reset')
    torch.cuda.current_stream(self.device).synchronize()

  def step(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """actions int32 [N] -> (reward f32 [N], done u8 [N], wind_uv f32 [N,2]); stream-ordered, async."""
    if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
      actions = actions.to(self.device, torch.int32).contiguous()
    rc = self._lib.ble_step_ex(self._h, _ptr(actions), ctypes.byref(self._step_out), self._stream())
    self._check(rc, 'ble_step_ex')
    return self._reward, self._done, self._wind

  def step_info(self) -> Dict[str, torch.Tensor]:
    """BalloonEnv.step's info (env/balloon_env.py:280-290) for the step that just ran, written by the step kernel:
    out_of_power / envelope_burst / zeropressure bool [N], time_elapsed int32 [N] (seconds), plus sim_error bool [N]
    (an atmosphere query or sunrise search of that balloon failed since its last reset; the reference raises)."""
    st = self._status
    return {'out_of_power': st == STATUS_OUT_OF_POWER, 'envelope_burst': st == STATUS_BURST,
            'zeropressure': st == STATUS_ZEROPRESSURE, 'time_elapsed': self._time_elapsed,
            'sim_error': self._sim_error != 0}

  def rollout(self, actions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """K consecutive steps in ONE launch for an open-loop action sequence: actions int32 [K, N] ->
    (reward f32 [K, N], done u8 [K, N]).  step_info() then describes the state after the last step."""
    if actions.dim() != 2 or actions.shape[1] != self.num_envs:
      raise ValueError('actions must be [K, N]')
    actions = actions.to(self.device, torch.int32).contiguous()
    k = int(actions.shape[0])
    reward = torch.empty(k, self.num_envs, dtype=torch.float32, device=self.device)
    done = torch.empty(k, self.num_envs, dtype=torch.uint8, device=self.device)
    out = _lib.BleStepOut(reward.data_ptr(), done.data_ptr(), self._wind.data_ptr(), self._status.data_ptr(),
                          self._time_elapsed.data_ptr(), self._sim_error.data_ptr())
    self._check(self._lib.ble_rollout(self._h, _ptr(actions), k, ctypes.byref(out), self._stream()), 'ble_rollout')
    return reward, done

  def step_host(self, actions: np.ndarray, reward: np.ndarray, done: np.ndarray):
    """Host-buffer step: int32 [N] in; float32 [N] / uint8 [N] out (copies inside, blocking)."""
    assert actions.dtype == np.int32 and reward.dtype == np.float32 and done.dtype == np.uint8
    rc = self._lib.ble_step_host(self._h, ctypes.c_void_p(actions.ctypes.data),
                                 ctypes.c_void_p(reward.ctypes.data), ctypes.c_void_p(done.ctypes.data),
                                 self._stream())
    self._check(rc, 'ble_step_host')


class Box:
  """Minimal stand-in for gym.spaces.Box (low / high / shape)."""

  def __init__(self, low: np.ndarray, high: np.ndarray):
    self.low, self.high, self.shape, self.dtype = low, high, low.shape, low.dtype

  def __repr__(self):
    return f'Box{self.shape}'


def perciatelli_observation_space() -> Box:
  """PerciatelliFeatureConstructor.observation_space (env/features.py:332-348)."""
  low = np.zeros(1099, np.float32)
  high = np.ones(1099, np.float32)
  low[[3, 4, 5, 6]] = -1.0          # sin / cos features
  low[15] = 1.0                      # ACS pressure ratio in [1, inf)
  high[15] = np.inf
  return Box(low, high)


class BatchedBalloonEnv:
  """Vectorised BalloonEnv: `step(actions[N]) -> (obs, reward[N], done[N], info)`.

  Reference: BalloonEnv (env/balloon_env.py:105-300).  `action_space.n == 3`.
  observation='perciatelli' returns the reference's default observation (float32 [N, 1099],
  PerciatelliFeatureConstructor, env/features.py:269-581) computed on the device;
  observation=None skips it (obs is None), which is the reference's hot path with a null feature
  constructor.
  """

  def __init__(self, num_envs: int, *, device: str = 'cuda:0', precision: str = 'fp32',
               wind_model: str = 'grid', enable_noise: bool = True, seed: int = 0,
               arena: Optional[BatchedBalloonArena] = None, observation: Optional[str] = None,
               field_layout: str = 'x64', decoder_params=None, shared_field_pool: Optional[torch.Tensor] = None,
               first_env: int = 0, broadcast_src: Optional[int] = 0, auto_reset: bool = False):
    """decoder_params (the flax tree of offlineskies22_decoder.msgpack['params']) switches the wind source
    to the reference's default GenerativeWindFieldSampler: every reset decodes one new field per balloon
    from that balloon's seed (env/generative_wind_field.py:52-62).  Without it the fields are whatever was
    loaded with arena.set_wind_fields / write_wind_fields, or `shared_field_pool` (float32 [F,21,21,10,9,2]):
    balloon i of the JOB flies field (first_env + i) % F.

    Under torch.distributed (one process per GPU, this rank holding the balloons [first_env, first_env + num_envs) of
    the job) the constructor is a collective: rank `broadcast_src` hands its decoder weights / field pool to the others
    over NCCL, so only one process reads them from disk (the others may pass None / an empty tensor of the right
    shape).  broadcast_src=None switches that off (every rank brings its own).

    auto_reset=True: a balloon whose step returned done starts a new episode before step() returns (reward / done /
    info describe the step that ended the episode, the observation is the new episode's first one), the vector-env
    convention; its seed continues a splitmix64 chain from the seed it was reset with.  With decoder_params the new
    episode also gets a new wind field (one device -> host read of the done count per step); without, the C library
    does it inside ble_step (ble_config.auto_reset) and the balloon keeps its field."""
    if observation not in (None, 'perciatelli'):
      raise ValueError("observation must be None or 'perciatelli'")
    self._auto_reset_host = bool(auto_reset) and decoder_params is not None      # new field per episode: driven from here
    self.arena = arena if arena is not None else BatchedBalloonArena(
        num_envs, device=device, precision=precision, wind_model=wind_model, enable_noise=enable_noise,
        field_layout=field_layout, enable_features=observation == 'perciatelli',
        auto_reset=bool(auto_reset) and decoder_params is None)
    if observation == 'perciatelli' and not self.arena.enable_features:
      raise ValueError("the arena was created without enable_features=True")
    self.observation = observation
    self.num_envs = self.arena.num_envs
    self.device = self.arena.device
    self._obs = (torch.empty(self.num_envs, 1099, dtype=torch.float32, device=self.device)
                 if observation == 'perciatelli' else None)
    self._generator = torch.Generator(device='cpu')
    self.seed(seed)
    if broadcast_src is not None and sharding.distributed_world() > 1:
      decoder_params = sharding.broadcast_decoder_params(decoder_params, self.device, src=broadcast_src)
      if shared_field_pool is not None:
        shared_field_pool = sharding.broadcast_field_pool(shared_field_pool.to(self.device, torch.float32).contiguous(),
                                                          src=broadcast_src)
    if shared_field_pool is not None:
      n_f = int(shared_field_pool.shape[0])
      env_to_field = (torch.arange(self.num_envs, dtype=torch.int64) + int(first_env)) % n_f
      self.arena.set_wind_fields(shared_field_pool, env_to_field.to(torch.int32))
    self._generative = decoder_params is not None
    if self._generative:
      self.arena.set_decoder(decoder_params)
      self.arena.alloc_wind_fields(self.num_envs)
      self.arena.set_field_map(torch.arange(self.num_envs, dtype=torch.int32, device=self.device))

  @property
  def action_space(self) -> Discrete:
    return Discrete(3)

  @property
  def observation_space(self) -> Optional[Box]:
    return perciatelli_observation_space() if self.observation == 'perciatelli' else None

  @property
  def reward_range(self) -> Tuple[float, float]:
    return (0.0, 1.0)

  def seed(self, seed: int) -> None:
    self._generator.manual_seed(int(seed))

  def reset(self, *, seed: Optional[int] = None, seeds: Optional[torch.Tensor] = None):
    """seed: reseeds the stream the per-balloon seeds are drawn from (BalloonEnv.seed + reset);
    seeds: int64 [N], one explicit episode seed per balloon (an evaluation suite's seeds)."""
    if seed is not None:
      self.seed(seed)
    if seeds is None:
      seeds = torch.randint(0, 2**62, (self.num_envs,), dtype=torch.int64, generator=self._generator)
    seeds = torch.as_tensor(seeds, dtype=torch.int64)
    if seeds.numel() != self.num_envs:
      raise ValueError('seeds must have one entry per balloon')
    if self._generative:
      self.arena.sample_wind_fields(seeds)
    self.arena.reset(seeds)
    return self._observe()

  def reset_where(self, mask: torch.Tensor):
    """Starts a new episode for the balloons with mask != 0 (the vectorised stand-in for the reference's
    per-environment `env.reset()` at an episode end); the others keep flying.  Returns the observation."""
    mask = mask.to(self.device).ne(0)
    seeds = torch.randint(0, 2**62, (self.num_envs,), dtype=torch.int64, generator=self._generator).to(self.device)
    if self._generative:
      idx = mask.nonzero().flatten()
      self.arena.sample_wind_fields_at(seeds[idx], idx.to(torch.int32))
    self.arena.reset(seeds, mask.to(torch.uint8))
    return self._observe()

  def _observe(self):
    if self.observation is None:
      return None
    return self.arena.features(self._obs)

  def step(self, actions: torch.Tensor):
    reward, done, _ = self.arena.step(actions)
    info = self.arena.step_info()
    if self._auto_reset_host and bool(done.any()):
      reward, done = reward.clone(), done.clone()          # the arena's output buffers are reused by the reset's observe
      return self.reset_where(done), reward, done, info
    return self._observe(), reward, done, info

  def get_info(self) -> Dict[str, torch.Tensor]:
    """_get_info (env/balloon_env.py:280-290) for all balloons, from the CURRENT state (also valid before the first
    step, e.g. right after reset); step() already returns the same dictionary for the state it produced."""
    st = self.arena.get_state_dict()
    status = st['status']
    return {'out_of_power': status == STATUS_OUT_OF_POWER, 'envelope_burst': status == STATUS_BURST,
            'zeropressure': status == STATUS_ZEROPRESSURE, 'time_elapsed': st['time_elapsed']}

  def close(self):
    self.arena.close()
