"""The three safety layers as batched integer state machines (ORACLE / test infrastructure).

Follows env/balloon/power_safety.py:33-126, env/balloon/envelope_safety.py:40-157 and
env/balloon/altitude_safety.py:35-111.  The reference drives the `transitions` library with
ordered first-match transition tables; the tables are restated here as explicit next-state
functions.
"""
import numpy as np

from oracle import constants as C


def paused_action(action):
  """DOWN -> STAY, otherwise unchanged (power_safety.py:120-126)."""
  return np.where(action == C.DOWN, C.STAY, action)


def power_safety_get_action(action, date_time, battery_charge_wh, sunrise_h, sunset, paused):
  """power_safety.py:52-118 -> (action, sunrise_h, sunset, paused).

  sunrise_h = sunrise + 30 min hysteresis (:50), all times int64 unix seconds.
  """
  day = C.NUM_SECONDS_PER_DAY
  # `while date_time > t: t += 1 day` (:83-86)
  k = np.where(date_time > sunrise_h, (date_time - sunrise_h + day - 1) // day, 0)
  sunrise_h = sunrise_h + k * day
  k = np.where(date_time > sunset, (date_time - sunset + day - 1) // day, 0)
  sunset = sunset + k * day

  is_day = sunset < sunrise_h                                              # :88
  soc = battery_charge_wh / C.BATTERY_CAPACITY_WH
  day_keep_paused = is_day & paused & (soc < C.POWER_SOC_RESTART)         # :92-93
  hours_to_sunrise = (sunrise_h - date_time) / 3600.0
  floating_charge = C.NIGHTTIME_POWER_LOAD_W * hours_to_sunrise            # :107-109
  expected = (battery_charge_wh - floating_charge) / C.BATTERY_CAPACITY_WH
  night_new_pause = (~is_day) & (~paused) & (expected < C.POWER_SOC_MIN)   # :111-115

  new_paused = np.where(is_day, day_keep_paused, paused | night_new_pause)
  out = np.where(new_paused, paused_action(action), action)
  return out, sunrise_h, sunset, new_paused


def envelope_safety_get_action(action, superpressure, state,
                               max_superpressure=C.ENVELOPE_MAX_SUPERPRESSURE):
  """envelope_safety.py:109-157 -> (action, state)."""
  sp = superpressure
  low_keep = (state == C.ENV_LOW_CRITICAL) | (state == C.ENV_LOW)
  high_keep = (state == C.ENV_HIGH) | (state == C.ENV_HIGH_CRITICAL)
  new_state = np.select(
      [sp < C.ENV_CRITICAL_BUFFER,
       sp < C.ENV_BUFFER,
       sp < C.ENV_BUFFER + C.ENV_RESTART_HYSTERESIS,
       sp < max_superpressure - C.ENV_BUFFER - C.ENV_RESTART_HYSTERESIS,
       sp < max_superpressure - C.ENV_BUFFER,
       sp < max_superpressure - C.ENV_CRITICAL_BUFFER],
      [C.ENV_LOW_CRITICAL,
       C.ENV_LOW,
       np.where(low_keep, C.ENV_LOW, C.ENV_NOMINAL),       # low_nominal   (:64-71)
       C.ENV_NOMINAL,
       np.where(high_keep, C.ENV_HIGH, C.ENV_NOMINAL),     # high_nominal  (:76-83)
       C.ENV_HIGH],
      default=C.ENV_HIGH_CRITICAL)
  critical = (new_state == C.ENV_LOW_CRITICAL) | (new_state == C.ENV_HIGH_CRITICAL)
  guarded = (new_state == C.ENV_LOW) | (new_state == C.ENV_HIGH)
  out = np.where(critical, C.UP, np.where(guarded, paused_action(action), action))
  return out, new_state


def altitude_safety_get_action(action, altitude_m, state):
  """altitude_safety.py:73-111 -> (action, state)."""
  was_low = (state == C.ALT_VERY_LOW) | (state == C.ALT_LOW)
  new_state = np.select(
      [altitude_m < C.ALT_MIN_ALTITUDE_M,
       altitude_m < C.ALT_MIN_ALTITUDE_M + C.ALT_BUFFER_M,
       altitude_m < C.ALT_MIN_ALTITUDE_M + C.ALT_BUFFER_M + C.ALT_RESTART_HYSTERESIS_M],
      [C.ALT_VERY_LOW,
       C.ALT_LOW,
       np.where(was_low, C.ALT_LOW, C.ALT_NOMINAL)],       # low_nominal (:51-58)
      default=C.ALT_NOMINAL)
  out = np.where(new_state == C.ALT_VERY_LOW, C.UP,
                 np.where(new_state == C.ALT_LOW, paused_action(action), action))
  return out, new_state
