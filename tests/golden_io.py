"""Helpers to load the committed golden fixtures (tests only)."""
import json
import os

import numpy as np

from oracle import atmosphere as atmosphere_lib
from oracle import balloon as balloon_lib
from oracle import env as env_lib
from oracle import wind as wind_lib
from tests.golden import fields as golden_fields

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_kat():
  with open(os.path.join(GOLDEN_DIR, 'kat.json')) as f:
    return json.load(f)


def load_traj():
  z = np.load(os.path.join(GOLDEN_DIR, 'traj.npz'))
  out = {}
  for name in z['names']:
    name = str(name)
    out[name] = {k.split('/', 1)[1]: z[k] for k in z.files if k.startswith(name + '/')}
  return out


def load_safety():
  """tests/golden/safety.npz: single reference steps from states in every safety band
  (tests/golden/tier0/make_safety_golden.py)."""
  z = np.load(os.path.join(GOLDEN_DIR, 'safety.npz'))
  return {k: z[k] for k in z.files}


def oracle_env_for_records(rec, float_fields, int_fields, sel=slice(None)):
  """One batched OracleEnv holding the pre-step states of single-step records (load_safety layout)."""
  b = batch_from_rows(float_fields, int_fields, rec['f'][sel], rec['i'][sel])
  arena = env_lib.OracleArena(
      b, atmosphere_lib.Atmosphere(rec['alpha'][sel]), fields=golden_fields.field_bank(),
      field_idx=np.maximum(rec['field'][sel], 0),
      noise=wind_lib.SimplexWindNoise(rec['seeds'][sel], rec['offsets'][sel]), static_wind=rec['field'][sel] < 0,
      power_safety_layer_enabled=rec['psl'][sel].astype(bool))
  return env_lib.OracleEnv(arena)


def batch_from_rows(float_fields, int_fields, f_rows, i_rows):
  """Rows [N, len(fields)] -> oracle BalloonBatch."""
  f_rows = np.atleast_2d(np.asarray(f_rows, np.float64))
  i_rows = np.atleast_2d(np.asarray(i_rows, np.int64))
  kw = {k: f_rows[:, j].copy() for j, k in enumerate(float_fields)}
  kw.update({k: i_rows[:, j].copy() for j, k in enumerate(int_fields)})
  return balloon_lib.BalloonBatch(**kw)


def oracle_env_for_scenarios(scs, float_fields, int_fields):
  """Builds one batched OracleEnv holding every scenario at its recorded initial state."""
  b = batch_from_rows(float_fields, int_fields, np.stack([sc['f0'] for sc in scs]),
                      np.stack([sc['i0'] for sc in scs]))
  atm = atmosphere_lib.Atmosphere([float(sc['alpha']) for sc in scs])
  noise = wind_lib.SimplexWindNoise(np.stack([sc['seeds'] for sc in scs]),
                                    np.stack([sc['offsets'] for sc in scs]))
  field = np.array([int(sc['field']) for sc in scs])
  arena = env_lib.OracleArena(
      b, atm, fields=golden_fields.field_bank(), field_idx=np.maximum(field, 0), noise=noise,
      static_wind=field < 0,
      power_safety_layer_enabled=np.array([bool(sc['power_safety']) for sc in scs]))
  return env_lib.OracleEnv(arena)


def oracle_env_for_scenario(sc, float_fields, int_fields):
  return oracle_env_for_scenarios([sc], float_fields, int_fields)
