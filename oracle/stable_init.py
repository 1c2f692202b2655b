"""Reset-time equilibrium solve (ORACLE / test infrastructure).  Follows env/balloon/stable_init.py:40-157."""
import numpy as np

from oracle import balloon as balloon_lib
from oracle import constants as C
from oracle import solar
from oracle import thermal


def calculate_stable_params_for_pressure(pressure, mols_lift_gas, lat, lng, date_time,
                                         upwelling_infrared, atmosphere):
  """-> dict(ambient_temperature, internal_temperature, mols_air, envelope_volume, superpressure)."""
  pressure = np.asarray(pressure, np.float64)
  _, ambient_temperature = atmosphere.at_pressure(pressure)                # :76
  mols_air = ((pressure * C.DRY_AIR_MOLAR_MASS * C.ENVELOPE_VOLUME_BASE /
               (C.UNIVERSAL_GAS_CONSTANT * ambient_temperature) -
               C.ENVELOPE_MASS - C.PAYLOAD_MASS - C.HE_MOLAR_MASS * mols_lift_gas)
              / C.DRY_AIR_MOLAR_MASS)                                      # :92-96
  mols_air = np.clip(mols_air, 0.0, None)                                  # :98
  internal_temperature = np.full(pressure.shape, 206.0)                    # :101
  el, _, flux = solar.solar_calculator(lat, lng, date_time)                # :102
  delta_temp = 0.01
  active = np.ones(pressure.shape, bool)
  for _ in range(10):                                                      # :107-127
    d1 = thermal.d_balloon_temperature_dt(
        C.ENVELOPE_VOLUME_BASE, C.ENVELOPE_MASS, internal_temperature - delta_temp / 2,
        ambient_temperature, pressure, el, flux, upwelling_infrared)
    d2 = thermal.d_balloon_temperature_dt(
        C.ENVELOPE_VOLUME_BASE, C.ENVELOPE_MASS, internal_temperature + delta_temp / 2,
        ambient_temperature, pressure, el, flux, upwelling_infrared)
    d2t = (d2 - d1) / delta_temp
    mean_d = (d1 + d2) / 2.0
    upd = active & (np.abs(d2t) > 0.0)
    with np.errstate(divide='ignore', invalid='ignore'):
      internal_temperature = np.where(upd, internal_temperature - mean_d / d2t,
                                      internal_temperature)
    active = active & ~(np.abs(mean_d) < 1e-5)
    if not active.any():
      break
  volume, superpressure = balloon_lib.calculate_superpressure_and_volume(
      mols_lift_gas, mols_air, internal_temperature, pressure)             # :130-133
  return dict(ambient_temperature=ambient_temperature, internal_temperature=internal_temperature,
              mols_air=mols_air, envelope_volume=volume, superpressure=superpressure)


def cold_start_to_stable_params(b, atmosphere):
  """In-place on a BalloonBatch; :139-157."""
  lat, lng = b.latlng()
  p = calculate_stable_params_for_pressure(
      b.pressure, b.mols_lift_gas, lat, lng, b.date_time, b.upwelling_infrared, atmosphere)
  for k, v in p.items():
    setattr(b, k, np.asarray(v, np.float64).copy())
