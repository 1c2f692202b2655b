"""Golden vectors for the evaluation surface, from the UNMODIFIED reference.

    python -m tests.golden.tier0.make_agent_golden

Writes tests/golden/agents.npz + tests/golden/eval.json:
  * open loop: the reference's StationSeekerAgent (agents/station_seeker_agent.py:72-178) on the
    recorded observations of features.npz and on synthetic 1099-feature vectors -> action, best
    level, 181 level scores;
  * closed loop: the reference's eval_lib.eval_agent (eval/eval_lib.py:123-211) flying the
    StationSeekerAgent in the reference's BalloonEnv from an injected initial state (env.seed /
    env.reset of that env INSTANCE are replaced by the harness so that the state, wind field and
    noise parameters are the injected ones; no reference file is edited);
  * the JSON produced by the reference's EvalResultEncoder for those results.

Scalar promotion: the reference pins numpy==1.19.5 (requirements.txt:40), where a float32 SCALAR
combined with a Python float gives float64, so StationSeeker's score arithmetic is fp64 on the
float32 features.  This container has NumPy 2 (NEP 50: the same expression stays float32).  The
harness therefore hands the agent the observation converted to float64 -- same values, and the
promotion of the pinned environment -- and records the float32 vector.
"""
import json
import os

import tests.golden.tier0.boot as boot  # noqa: F401
import tests.golden.tier0.make_golden as mg

import jax
import numpy as np
from balloon_learning_environment.agents import station_seeker_agent
from balloon_learning_environment.env import balloon_arena
from balloon_learning_environment.env import balloon_env
from balloon_learning_environment.env import features as features_lib
from balloon_learning_environment.env import grid_based_wind_field
from balloon_learning_environment.env import wind_field
from balloon_learning_environment.eval import eval_lib
from balloon_learning_environment.eval import suites
from balloon_learning_environment.utils import test_helpers
from balloon_learning_environment.utils import units

from tests.golden import fields as golden_fields

OUT = mg.OUT


def synthetic_observations(rng, count):
  """Feature vectors with the structure PerciatelliFeatureConstructor emits: a contiguous band of
  valid levels, the rest flagged invalid as (uncertainty 0, bearing 1, magnitude 1)."""
  obs = np.zeros((count, 1099), np.float32)
  for k in range(count):
    amb = rng.uniform(0, 1, 16).astype(np.float32)
    amb[7] = np.float32(rng.uniform(0, 0.92))                        # distance squash d / (d + 250 km)
    obs[k, :16] = amb
    w = np.zeros((361, 3), np.float32)
    w[:, 0] = 0.0; w[:, 1] = 1.0; w[:, 2] = 1.0
    lo = int(rng.integers(60, 180)); hi = int(rng.integers(181, 300))
    w[lo:hi, 0] = rng.uniform(0, 1, hi - lo)
    w[lo:hi, 1] = rng.uniform(0, 1, hi - lo)
    w[lo:hi, 2] = rng.uniform(0, 0.6, hi - lo)
    if k % 7 == 0:                                                   # plateaus: exact ties between levels
      w[lo:hi] = w[lo]
    obs[k, 16:] = w.reshape(-1)
  return obs


def open_loop(agent, obs):
  actions, best, scores = [], [], []
  for o in obs:
    o = np.asarray(o, np.float64)                       # numpy-1.19 scalar promotion, see module docstring
    named = features_lib.NamedPerciatelliFeatures(o)
    level, sc = agent.find_best_pressure_level(named)
    actions.append(int(agent.pick_action(o))); best.append(int(level)); scores.append(np.asarray(sc, np.float64))
  return np.asarray(actions, np.int64), np.asarray(best, np.int64), np.asarray(scores)


def closed_loop(sc, bank, rng, steps, seed_label):
  date = units.datetime(*sc['date'])
  atm = mg.make_atmosphere(sc['alpha'])
  b = test_helpers.create_balloon(
      x=units.Distance(m=sc['x']), y=units.Distance(m=sc['y']), center_lat=sc['lat'], center_lng=sc['lng'],
      pressure=sc['pressure'], power_percent=sc['power'], date_time=date, upwelling_infrared=sc['ir'],
      atmosphere=atm)
  wf = grid_based_wind_field.GridBasedWindField(mg._BankSampler(bank[sc['field']]))
  wf.reset(jax.random.PRNGKey(1), date)
  seeds = rng.integers(0, 1634753849, size=(2, 5))
  offsets = (rng.uniform(0, 1, size=(2, 5, 4)).astype(np.float32) * np.float32(2.0) - np.float32(1.0)).astype(np.float64)
  arena = balloon_arena.BalloonArena(features_lib.PerciatelliFeatureConstructor, wf, seed=0)
  env = balloon_env.BalloonEnv(arena=arena, seed=0)
  atm = mg.make_atmosphere(sc['alpha'])
  mg._inject_noise(wf, seeds, offsets)
  wf.field = bank[sc['field']]
  arena._balloon = b
  arena._atmosphere = atm
  arena.feature_constructor = features_lib.PerciatelliFeatureConstructor(wf, atm)
  arena.feature_constructor.observe(arena.get_measurements())
  obs0 = arena.feature_constructor.get_features()
  f0, i0 = mg.snapshot(b.state)

  log = dict(obs=[obs0], actions=[], reward=[])
  real_step = env.step

  def logged_step(action):
    out = real_step(action)
    log['actions'].append(int(action)); log['obs'].append(out[0]); log['reward'].append(float(out[1]))
    return (np.asarray(out[0], np.float64),) + tuple(out[1:])       # numpy-1.19 scalar promotion

  env.seed = lambda s: None                 # the injected state IS the episode of this "seed"
  env.reset = lambda: np.asarray(obs0, np.float64)
  env.step = logged_step
  agent = station_seeker_agent.StationSeekerAgent(3, (1099,))
  results = eval_lib.eval_agent(agent, env, suites.EvaluationSuite([seed_label], steps))
  res = results[0]
  path = np.asarray([[p.x.kilometers, p.y.kilometers, p.pressure, p.superpressure,
                      p.time_elapsed.total_seconds(), p.battery_soc] for p in res.flight_path])
  out = dict(alpha=sc['alpha'], field=sc['field'], power_safety=1, seeds=seeds, offsets=offsets,
             f0=np.asarray(f0), i0=np.asarray(i0, np.int64), actions=np.asarray(log['actions'], np.int64),
             reward=np.asarray(log['reward']), obs=np.asarray(log['obs'], np.float32), flight_path=path,
             cumulative_reward=float(res.cumulative_reward), time_within_radius=float(res.time_within_radius),
             final_timestep=int(res.final_timestep),
             flags=np.asarray([res.out_of_power, res.envelope_burst, res.zeropressure], np.int64))
  return out, results


def main():
  agent = station_seeker_agent.StationSeekerAgent(3, (1099,))
  rec = np.load(os.path.join(OUT, 'features.npz'))
  recorded = np.concatenate([rec[f'{n}/obs'] for n in rec['names']])
  synth = synthetic_observations(np.random.default_rng(77), 256)
  obs = np.concatenate([recorded, synth]).astype(np.float32)
  actions, best, scores = open_loop(agent, obs)
  flat = dict(open_obs=obs, open_actions=actions, open_best=best, open_scores=scores,
              n_recorded=np.int64(len(recorded)))

  bank = golden_fields.field_bank()
  rng = np.random.default_rng(515)
  scenarios = [
      dict(name='seeker_near', alpha=0.35, field=1, x=-40e3, y=25e3, lat=2.0, lng=-50.0, pressure=9300.0, power=0.95,
           date=(2013, 6, 1, 4, 0, 0), ir=250.0),
      dict(name='seeker_far', alpha=0.8, field=2, x=160e3, y=-120e3, lat=-6.0, lng=100.0, pressure=7800.0, power=0.8,
           date=(2012, 11, 20, 15, 30, 0), ir=300.0),
  ]
  names, all_results = [], []
  for k, sc in enumerate(scenarios):
    out, results = closed_loop(sc, bank, rng, 100, seed_label=k)
    names.append(sc['name']); all_results += results
    for key, v in out.items():
      flat[f"{sc['name']}/{key}"] = np.asarray(v)
    print(sc['name'], 'actions', np.bincount(out['actions'], minlength=3), 'twr', out['time_within_radius'],
          'cum reward', out['cumulative_reward'])
  flat['names'] = np.asarray(names)
  np.savez_compressed(os.path.join(OUT, 'agents.npz'), **flat)
  with open(os.path.join(OUT, 'eval.json'), 'w') as f:
    f.write(json.dumps(all_results, cls=eval_lib.EvalResultEncoder))
  print('open loop', len(obs), 'observations; action histogram', np.bincount(actions, minlength=3))


if __name__ == '__main__':
  main()
