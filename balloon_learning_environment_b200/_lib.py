"""ctypes binding of the C ABI in include/ble_b200.h.

There is no CPU fallback: if the shared library is missing or no CUDA device is present the
import / create call raises.
"""
import ctypes
import os

from balloon_learning_environment_b200 import _build

_c = ctypes
_LIB = None

BLE_OK = 0
PRECISION = {'fp32': 0, 'fp64': 1}
WIND_MODEL = {'grid': 0, 'simple_static': 1}
FIELD_LAYOUT = {'x64': 0, 'x128': 1}

# Row order of the state exchange matrices (include/ble_b200.h, BLE_F_* / BLE_I_*).
F_ROWS = ('x', 'y', 'pressure', 'ambient_temperature', 'internal_temperature', 'envelope_volume',
          'superpressure', 'mols_air', 'mols_lift_gas', 'battery_charge', 'acs_power', 'acs_mass_flow',
          'solar_charging', 'power_load', 'center_lat', 'center_lng', 'upwelling_infrared',
          'atmosphere_alpha')
I_ROWS = ('date_time', 'time_elapsed', 'last_command', 'status', 'envelope_state', 'altitude_state',
          'power_paused', 'sunrise_h', 'sunset', 'power_safety_enabled')

D_ROWS = ('lat', 'lng', 'solar_elevation', 'solar_flux', 'excess_energy', 'navigation_is_paused',
          'pressure_ratio', 'battery_soc', 'altitude')

E_ROWS = ('cumulative_reward', 'time_within_radius', 'out_of_power', 'envelope_burst', 'zeropressure',
          'final_timestep', 'active')           # BLE_E_* of include/ble_b200.h
EXPORTS = ('ble_create', 'ble_destroy', 'ble_last_error', 'ble_num_envs', 'ble_upload_fields',
           'ble_alloc_fields', 'ble_write_fields', 'ble_set_field_map', 'ble_set_decoder', 'ble_decode_fields',
           'ble_set_noise', 'ble_state_upload', 'ble_state_download', 'ble_reset', 'ble_init_derived',
           'ble_step', 'ble_step_ex', 'ble_rollout', 'ble_step_host', 'ble_wind_at_balloon', 'ble_wind_gather', 'ble_wind_query', 'ble_atmosphere_query', 'ble_derived',
           'ble_features_perciatelli', 'ble_features_observe', 'ble_features_clear', 'ble_features_track',
           'ble_generate_fields', 'ble_generate_fields_at', 'ble_sample_latents', 'ble_agent_station_seeker', 'ble_agent_random_walk',
           'ble_eval_begin', 'ble_eval_accumulate', 'ble_eval_results',
           'ble_launch_count',
           'ble_qr_greedy', 'ble_qr_target', 'ble_qr_loss', 'ble_replay_sample', 'ble_adam_step',
           'ble_marco_polo_step', 'ble_dense_tf32', 'ble_transpose_f32', 'ble_row_sum_f32')


class BleConfig(_c.Structure):
  _fields_ = [('precision', _c.c_int32), ('wind_model', _c.c_int32), ('enable_noise', _c.c_int32),
              ('field_layout', _c.c_int32), ('enable_features', _c.c_int32), ('decoder_fp32', _c.c_int32), ('auto_reset', _c.c_int32),
              ('reserved', _c.c_int32 * 1)]


class BleReplayView(_c.Structure):
  _fields_ = [('obs', _c.c_void_p), ('action', _c.c_void_p), ('reward', _c.c_void_p), ('terminal', _c.c_void_p),
              ('truncated', _c.c_void_p), ('capacity', _c.c_int64), ('num_envs', _c.c_int64), ('count', _c.c_int64),
              ('n_step', _c.c_int32), ('num_features', _c.c_int32), ('gamma', _c.c_float), ('out_pitch', _c.c_int32)]


class BleStepOut(_c.Structure):
  _fields_ = [('reward', _c.c_void_p), ('done', _c.c_void_p), ('wind_uv', _c.c_void_p), ('status', _c.c_void_p),
              ('time_elapsed', _c.c_void_p), ('sim_error', _c.c_void_p), ('reserved', _c.c_void_p * 2)]


class BleStateSoa(_c.Structure):
  _fields_ = [('f64', _c.c_void_p), ('i64', _c.c_void_p)]


class BleError(RuntimeError):
  pass


def library_path():
  return _build.LIB_PATH


def load(build_if_missing=True):
  """Loads libble_b200.so (building it in-tree first if it is missing and nvcc is present)."""
  global _LIB
  if _LIB is not None:
    return _LIB
  path = library_path()
  if not os.path.exists(path):
    if not build_if_missing:
      raise BleError(f'{path} is missing; run `python -m balloon_learning_environment_b200._build`')
    _build.build()
  lib = _c.CDLL(path)
  vp, i64, i32 = _c.c_void_p, _c.c_int64, _c.c_int32
  lib.ble_create.argtypes = [_c.c_int, i64, _c.POINTER(BleConfig), _c.POINTER(vp)]
  lib.ble_destroy.argtypes = [vp]
  lib.ble_last_error.argtypes = [vp]
  lib.ble_last_error.restype = _c.c_char_p
  lib.ble_num_envs.argtypes = [vp]
  lib.ble_num_envs.restype = i64
  lib.ble_launch_count.argtypes = [vp]
  lib.ble_launch_count.restype = i64
  lib.ble_upload_fields.argtypes = [vp, vp, i64, vp, vp]
  lib.ble_alloc_fields.argtypes = [vp, i64, vp]
  lib.ble_write_fields.argtypes = [vp, vp, i64, i64, vp]
  lib.ble_set_field_map.argtypes = [vp, vp, vp]
  lib.ble_set_noise.argtypes = [vp, vp, vp, vp]
  lib.ble_set_decoder.argtypes = [vp, _c.POINTER(vp), _c.POINTER(vp), vp]
  lib.ble_decode_fields.argtypes = [vp, vp, i64, vp, vp]
  lib.ble_state_upload.argtypes = [vp, _c.POINTER(BleStateSoa), vp]
  lib.ble_state_download.argtypes = [vp, _c.POINTER(BleStateSoa), vp]
  lib.ble_reset.argtypes = [vp, vp, vp, vp]
  lib.ble_init_derived.argtypes = [vp, i32, vp]
  lib.ble_step.argtypes = [vp, vp, vp, vp, vp, vp]
  lib.ble_step_ex.argtypes = [vp, vp, _c.POINTER(BleStepOut), vp]
  lib.ble_rollout.argtypes = [vp, vp, i32, _c.POINTER(BleStepOut), vp]
  lib.ble_step_host.argtypes = [vp, vp, vp, vp, vp]
  lib.ble_wind_at_balloon.argtypes = [vp, vp, vp]
  lib.ble_wind_gather.argtypes = [vp, vp, vp, vp, i64, vp]
  lib.ble_wind_query.argtypes = [vp, vp, vp, i32, vp, i64, vp]
  lib.ble_atmosphere_query.argtypes = [vp, i32, vp, vp, vp, i64, vp]
  lib.ble_derived.argtypes = [vp, vp, vp]
  lib.ble_features_perciatelli.argtypes = [vp, vp, vp]
  lib.ble_features_observe.argtypes = [vp, vp]
  lib.ble_features_clear.argtypes = [vp, vp]
  lib.ble_features_track.argtypes = [vp, i32]
  lib.ble_generate_fields.argtypes = [vp, vp, i64, i64, vp]
  lib.ble_generate_fields_at.argtypes = [vp, vp, vp, i64, vp]
  lib.ble_sample_latents.argtypes = [vp, vp, i64, vp, vp]
  lib.ble_agent_station_seeker.argtypes = [vp, vp, vp, vp, vp]
  lib.ble_agent_random_walk.argtypes = [vp, vp, vp, i32, vp, vp]
  lib.ble_eval_begin.argtypes = [vp, vp]
  lib.ble_eval_accumulate.argtypes = [vp, vp, vp, vp]
  lib.ble_eval_results.argtypes = [vp, vp, vp]
  f32, f64, u64 = _c.c_float, _c.c_double, _c.c_uint64
  lib.ble_qr_greedy.argtypes = [vp, i64, i32, i32, vp, vp, vp]
  lib.ble_qr_target.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp]
  lib.ble_qr_loss.argtypes = [vp, vp, vp, vp, f32, i64, i32, i32, f32, vp, vp, vp]
  lib.ble_replay_sample.argtypes = [_c.POINTER(BleReplayView), vp, u64, i64, vp, vp, vp, vp, vp, vp, vp, vp]
  lib.ble_adam_step.argtypes = [vp, vp, vp, vp, i64, f64, f64, f64, f64, i64, f32, vp]
  lib.ble_marco_polo_step.argtypes = [vp, vp, vp, i64, vp, vp, vp, i64, f32, vp, vp]
  lib.ble_dense_tf32.argtypes = [vp, i64, vp, i64, i64, i64, i64, i32, vp, i64, vp, i64, vp, i64, i32, vp, i64, vp]
  lib.ble_transpose_f32.argtypes = [vp, i64, i64, i64, vp, i64, vp]
  lib.ble_row_sum_f32.argtypes = [vp, i64, i64, i64, vp, i32, vp]
  for name in EXPORTS:
    if name not in ('ble_last_error', 'ble_num_envs', 'ble_launch_count'):
      getattr(lib, name).restype = _c.c_int
  _LIB = lib
  return lib


def check(lib, handle, rc, what):
  if rc != BLE_OK:
    msg = lib.ble_last_error(handle)
    raise BleError(f'{what} failed (code {rc}): {msg.decode() if msg else "?"}')
