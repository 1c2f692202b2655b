"""Baseline controllers and the evaluation loop (ORACLE / test infrastructure, NumPy fp64).

Follows agents/station_seeker_agent.py:37-178 (StationSeeker score and action rule),
agents/random_walk_agent.py:35-94 (the deterministic part: hysteresis band around a target
pressure), env/features.py:148-266 (feature naming, validity, de-normalisation),
utils/transforms.py:45-94 and eval/eval_lib.py:33-211 (evaluation loop and result schema).
Vectorised over N observations; only tests/, smoke() and bench.py's CPU legs may import this.
"""
import numpy as np

from oracle import constants as C

NUM_LEVELS = 361                       # 2 * 181 - 1 relative levels (features.py:291)
CENTER = NUM_LEVELS // 2               # wind_column_center (features.py:253-256)

# agents/station_seeker_agent.py:44-57
HALF_RADIUS = 35.0
MAGNITUDE_WEIGHT = 0.07
CLOSE_BEARING_WEIGHT = 0.6
FAR_BEARING_WEIGHT = 0.45
CLOSE_BEARING = 250.0
FAR_BEARING = 500.0
DEFAULT_SCORE = 0.5
HYSTERESIS_K2 = 0.05
HYSTERESIS_K3 = 0.001
CONFIDENCE_EPSILON = 0.01


def wind_columns(obs):
  """[N,1099] -> (uncertainty, bearing, magnitude) each [N,361] as float64, and validity."""
  w = np.asarray(obs, np.float32)[:, 16:].reshape(-1, NUM_LEVELS, 3)
  unc, bearing, mag = (w[:, :, k] for k in range(3))
  valid = (mag != np.float32(1.0)) | (bearing != np.float32(1.0)) | (unc != np.float32(0.0))   # features.py:155-160
  return unc.astype(np.float64), bearing.astype(np.float64), mag.astype(np.float64), valid


def station_seeker_scores(obs):
  """altitude_score for every level (station_seeker_agent.py:115-178); invalid levels score 0."""
  obs = np.asarray(obs, np.float32)
  unc, bearing_n, mag_n, valid = wind_columns(obs)
  bearing = bearing_n * np.pi                                     # undo_linear_rescale (features.py:264-265)
  with np.errstate(divide='ignore', invalid='ignore'):
    magnitude = mag_n * 30.0 / (1.0 - mag_n)                      # undo_squash (transforms.py:88-94)
    d = obs[:, 7].astype(np.float64)
    distance = (d * 250.0 / (1.0 - d))[:, None]
    coeff = np.clip((distance - CLOSE_BEARING) / (FAR_BEARING - CLOSE_BEARING), 0.0, 1.0)
    bearing_weight = CLOSE_BEARING_WEIGHT + coeff * (FAR_BEARING_WEIGHT - CLOSE_BEARING_WEIGHT)
    alpha_delta = np.exp(-distance / HALF_RADIUS)
    wind_score = ((1 - alpha_delta) * np.exp(-bearing_weight * bearing)
                  + alpha_delta * np.exp(-MAGNITUDE_WEIGHT * magnitude))
    level_distance = np.abs(np.arange(NUM_LEVELS) - CENTER)[None, :]
    hysteresis = HYSTERESIS_K2 * np.exp(-HYSTERESIS_K3 * level_distance)
    score = (1.0 - unc + CONFIDENCE_EPSILON) * wind_score + unc * DEFAULT_SCORE + hysteresis
  return np.where(valid, score, 0.0)


def station_seeker_actions(obs):
  """pick_action (:72-88): first level with the strictly largest score; above centre -> DOWN."""
  scores = station_seeker_scores(obs)
  best = np.argmax(scores, axis=1)                                # first maximum == the reference's strict '>' scan
  if np.any(scores[np.arange(len(best)), best] <= 0.0):
    raise AssertionError('at least one pressure level should be valid')     # :109-110
  action = np.where(best < CENTER, C.UP, np.where(best > CENTER, C.DOWN, C.STAY))
  return action.astype(np.int64), best, scores


def balloon_pressure_from_obs(obs):
  """NamedPerciatelliFeatures.balloon_pressure (features.py:192-195), float32 feature -> Pa."""
  return np.asarray(obs, np.float32)[:, 0].astype(np.float64) * (C.PERCIATELLI_PRESSURE_RANGE_MAX - C.PERCIATELLI_PRESSURE_RANGE_MIN) \
      + C.PERCIATELLI_PRESSURE_RANGE_MIN


def random_walk_actions(obs, target_pressure, hysteresis=100.0):
  """RandomWalkAgent._select_action (random_walk_agent.py:62-73)."""
  p = balloon_pressure_from_obs(obs)
  t = np.asarray(target_pressure, np.float64)
  return np.where(p - hysteresis > t, C.UP, np.where(p + hysteresis < t, C.DOWN, C.STAY)).astype(np.int64)


def eval_agent(policy, env, feat, max_episode_length, radius_km=C.REWARD_RADIUS_KM):
  """eval_lib.eval_agent (:123-211) for a batch: every balloon is one seed's episode.

  policy(obs[N,1099]) -> actions[N]; env: oracle OracleEnv; feat: oracle PerciatelliFeatures already
  holding the first observation.  Returns the per-balloon EvaluationResult fields as arrays."""
  s = env.arena.state
  n = len(s.x)
  total = np.zeros(n); within = np.zeros(n, np.int64); steps = np.zeros(n, np.int64)
  active = s.status == C.STATUS_OK
  path = []
  action = policy(feat.get_features())
  for _ in range(max_episode_length):
    if not active.any():
      break
    reward, done, _ = env.step(np.where(active, action, C.STAY))
    feat.observe(mask=active)
    total += np.where(active, reward, 0.0)
    within += active & (np.sqrt(s.x * s.x + s.y * s.y) <= radius_km * 1000.0)          # :119-121
    steps += active
    path.append(np.stack([s.x / 1000.0, s.y / 1000.0, s.pressure, s.superpressure,
                          s.time_elapsed.astype(np.float64), s.battery_charge / C.BATTERY_CAPACITY_WH], axis=1))
    active = active & ~done
    action = policy(feat.get_features())
  return dict(cumulative_reward=total, time_within_radius=within / np.maximum(steps, 1), final_timestep=steps,
              out_of_power=s.status == C.STATUS_OUT_OF_POWER, envelope_burst=s.status == C.STATUS_BURST,
              zeropressure=s.status == C.STATUS_ZEROPRESSURE, flight_path=np.asarray(path))
