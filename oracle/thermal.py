"""Envelope heat balance (ORACLE / test infrastructure).  Follows env/balloon/thermal.py:28-230."""
import numpy as np

from oracle import solar

_SOLAR_VIEW_FACTOR = 0.25
_EARTH_VIEW_FACTOR = 0.4605
_REFLECTIVITY = 0.0291
_ABS_SOLAR = 0.01435
_ABS_IR_BASE = 0.04587
_ABS_IR_D_T = 0.000232
_ABS_IR_REF_T = 210
_FILM_SPECIFIC_HEAT = 1500
_STEFAN_BOLTZMAN = 0.000000056704
_R = 8.3144621
_M_AIR = 0.028964922481160


def absorptivity_ir(t_k):                                                # :77-91
  return _ABS_IR_BASE + _ABS_IR_D_T * (t_k - _ABS_IR_REF_T)


def total_absorptivity(absorptivity, reflectivity=_REFLECTIVITY):         # :92-147
  transmissivity = 1.0 - absorptivity - reflectivity
  factor = absorptivity * (1.0 + transmissivity / (1.0 - reflectivity))
  if np.any(factor < 0.0) or np.any(factor > 1.0):
    raise ValueError('total_absorptivity out of range')
  return factor


def convective_heat_air_factor(radius, t_balloon, t_ambient, pressure):   # :150-172
  viscosity = 1.458e-6 * (t_ambient ** 1.5) / (t_ambient + 110.4)
  conductivity = 0.0241 * ((t_ambient / 273.15) ** 0.9)
  prandtl = 0.804 - 3.25e-4 * t_ambient
  air_density = pressure * _M_AIR / (_R * t_ambient)
  grashof = (9.80665 * (air_density ** 2) * ((2 * radius) ** 3) /
             (t_ambient * (viscosity ** 2))) * np.abs(t_ambient - t_balloon)
  rayleigh = prandtl * grashof
  nusselt = 2 + 0.457 * (rayleigh ** 0.25) + ((1 + 2.69e-8 * rayleigh) ** (1.0 / 12.0))
  k_heat_transfer = nusselt * conductivity / (2 * radius)
  return k_heat_transfer * (t_ambient - t_balloon)


def d_balloon_temperature_dt(volume, mass, t_balloon, t_ambient, pressure,
                             solar_elevation_deg, solar_flux, earth_flux):  # :175-230
  radius = (3 * volume / (4 * np.pi)) ** (1 / 3)
  area = 4 * np.pi * radius * radius
  att = solar.solar_atmospheric_attenuation(solar_elevation_deg, pressure)
  q_solar = solar_flux * att * _SOLAR_VIEW_FACTOR * area * total_absorptivity(_ABS_SOLAR)
  q_earth = earth_flux * _EARTH_VIEW_FACTOR * area * total_absorptivity(
      absorptivity_ir((earth_flux / _STEFAN_BOLTZMAN) ** 0.25))
  q_emitted = (_STEFAN_BOLTZMAN * t_balloon ** 4) * area * total_absorptivity(
      absorptivity_ir(t_balloon))
  q_convective = area * convective_heat_air_factor(radius, t_balloon, t_ambient, pressure)
  return (q_solar + q_earth + q_convective - q_emitted) / (_FILM_SPECIFIC_HEAT * mass)
