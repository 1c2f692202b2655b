"""Evaluation surface: the oracle's controllers and evaluation loop against
the reference's StationSeekerAgent / eval_lib.eval_agent outputs (tests/golden/agents.npz, eval.json)."""
import json
import os

import numpy as np

from oracle import agents, features as F
from tests import golden_io

KAT = golden_io.load_kat()
FF, IF = KAT['float_fields'], KAT['int_fields']
GOLD = np.load(os.path.join(golden_io.GOLDEN_DIR, 'agents.npz'))


def scenario(name):
  return {k.split('/', 1)[1]: GOLD[k] for k in GOLD.files if k.startswith(name + '/')}


def test_station_seeker_open_loop_matches_reference():
  """499 observations (242 recorded, 257 synthetic incl. exact ties): action and best level bit-exact,
  all 361 level scores to 1e-13."""
  actions, best, scores = agents.station_seeker_actions(GOLD['open_obs'])
  np.testing.assert_array_equal(actions, GOLD['open_actions'])
  np.testing.assert_array_equal(best, GOLD['open_best'])
  np.testing.assert_allclose(scores, GOLD['open_scores'], rtol=1e-13, atol=1e-15)


def test_eval_loop_matches_reference_eval_agent():
  """The reference's eval_agent flying StationSeeker for 100 steps from an injected state: the oracle's
  closed loop (own features, own agent, own physics) picks the same actions and reports the same result."""
  for name in GOLD['names']:
    sc = scenario(str(name))
    env = golden_io.oracle_env_for_scenario(sc, FF, IF)
    feat = F.PerciatelliFeatures(env.arena)
    feat.observe()
    taken = []

    def policy(obs):
      a = agents.station_seeker_actions(obs)[0]
      taken.append(int(a[0]))
      return a

    res = agents.eval_agent(policy, env, feat, max_episode_length=len(sc['actions']))
    np.testing.assert_array_equal(taken[:len(sc['actions'])], sc['actions'])
    np.testing.assert_allclose(res['cumulative_reward'][0], sc['cumulative_reward'], rtol=1e-9)
    assert res['time_within_radius'][0] == sc['time_within_radius']
    assert res['final_timestep'][0] == sc['final_timestep']
    np.testing.assert_allclose(res['flight_path'][:, 0, :], sc['flight_path'], rtol=1e-8, atol=1e-9)


def test_reference_eval_json_schema():
  """eval/eval_lib.py:33-56: keys of an encoded EvaluationResult and of a flight-path point."""
  with open(os.path.join(golden_io.GOLDEN_DIR, 'eval.json')) as f:
    ref = json.load(f)
  assert list(ref[0].keys()) == ['seed', 'cumulative_reward', 'time_within_radius', 'out_of_power',
                                 'envelope_burst', 'zeropressure', 'final_timestep', 'flight_path']
  assert list(ref[0]['flight_path'][0].keys()) == ['x', 'y', 'pressure', 'superpressure', 'elapsed_seconds', 'power']


def test_device_agent_rules_replayed_on_host():
  """ble_agents.cuh compiled with g++: StationSeeker action / best level bit-exact against the reference on the
  499 golden observations, scores to 1e-13; RandomWalk band rule against the oracle."""
  import ctypes
  from tests import hostemu
  lib = hostemu.load()
  obs = np.ascontiguousarray(GOLD['open_obs'], np.float32)
  n = len(obs)
  actions = np.zeros(n, np.int32); best = np.zeros(n, np.int32); scores = np.zeros((n, 361))
  lib.emu_station_seeker(ctypes.c_int64(n), hostemu.ptr(obs), hostemu.ptr(actions), hostemu.ptr(best), hostemu.ptr(scores))
  np.testing.assert_array_equal(actions, GOLD['open_actions'])
  np.testing.assert_array_equal(best, GOLD['open_best'])
  np.testing.assert_allclose(scores, GOLD['open_scores'], rtol=1e-13, atol=1e-15)
  rng = np.random.default_rng(5)
  target = rng.uniform(5000, 14000, n)
  lib.emu_random_walk(ctypes.c_int64(n), hostemu.ptr(obs), hostemu.ptr(target), hostemu.ptr(actions))
  np.testing.assert_array_equal(actions, agents.random_walk_actions(obs, target))
