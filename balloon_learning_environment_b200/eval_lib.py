"""Vectorised evaluation (eval/eval_lib.py:33-211): every seed of a suite is one balloon of the batch.

The loop body (reward sum, steps within radius, flight path, stop at a terminal state) runs in
k_eval_accumulate; this module only sequences agent.step / env.step and shapes the results into the
reference's EvaluationResult / JSON schema.
"""
import dataclasses
import json
from typing import Any, List, Sequence

import numpy as np
import torch

from balloon_learning_environment_b200 import agents as agents_lib
from balloon_learning_environment_b200 import suites


@dataclasses.dataclass
class SimpleBalloonState:            # eval/eval_lib.py:58-80 (plain floats instead of unit objects)
  x_km: float
  y_km: float
  pressure: float
  superpressure: float
  elapsed_seconds: float
  battery_soc: float


@dataclasses.dataclass
class EvaluationResult:              # eval/eval_lib.py:84-116
  seed: int
  cumulative_reward: float
  time_within_radius: float
  out_of_power: bool
  envelope_burst: bool
  zeropressure: bool
  final_timestep: int
  flight_path: Sequence[SimpleBalloonState]

  def __str__(self) -> str:
    return (f'EvaluationResult(seed={self.seed}, cumulative_reward={self.cumulative_reward}, '
            f'time_within_radius={self.time_within_radius}, out_of_power={self.out_of_power}, '
            f'final_timestep={self.final_timestep})')


class EvalResultEncoder(json.JSONEncoder):
  """eval/eval_lib.py:33-56: same keys, same nesting."""

  def default(self, o: Any):
    if isinstance(o, SimpleBalloonState):
      return {'x': o.x_km, 'y': o.y_km, 'pressure': o.pressure, 'superpressure': o.superpressure,
              'elapsed_seconds': o.elapsed_seconds, 'power': o.battery_soc}
    if dataclasses.is_dataclass(o):
      return o.__dict__
    if isinstance(o, torch.Tensor) and o.numel() == 1:
      return o.item()
    if isinstance(o, (np.ndarray, np.generic)) and o.size == 1:
      return o.item()
    return json.JSONEncoder.default(self, o)


def eval_agent(agent: agents_lib.BatchedAgent, env, eval_suite: suites.EvaluationSuite, *,
               calculate_flight_path: bool = True) -> List[EvaluationResult]:
  """Flies every seed of the suite at once.  env: BatchedBalloonEnv(observation='perciatelli') with
  num_envs == len(eval_suite.seeds)."""
  assert eval_suite.max_episode_length > 0, 'max_episode_length must be > 0.'
  n = env.num_envs
  if n != len(eval_suite.seeds):
    raise ValueError(f'the env holds {n} balloons but the suite has {len(eval_suite.seeds)} seeds')
  arena = env.arena
  agent.set_mode(agents_lib.AgentMode.EVAL)
  observation = env.reset(seeds=torch.as_tensor(list(eval_suite.seeds), dtype=torch.int64))
  arena.eval_begin()
  action = agent.begin_episode(observation)
  path = (torch.empty(eval_suite.max_episode_length, 6, n, dtype=torch.float32, device=env.device)
          if calculate_flight_path else None)
  flown = 0
  reward = done = None
  for t in range(eval_suite.max_episode_length):
    observation, reward, done, _ = env.step(action)
    arena.eval_accumulate(reward, path[t] if path is not None else None)
    action = agent.step(reward, observation)
    flown = t + 1
    if t % 32 == 31 and not bool(arena.eval_results()['active'].any()):       # every flight has ended
      break
  agent.end_episode(reward, done)
  res = {k: v.cpu().numpy() for k, v in arena.eval_results().items()}
  path_h = path[:flown].cpu().numpy() if path is not None else None
  results = []
  for e, seed in enumerate(eval_suite.seeds):
    steps = int(res['final_timestep'][e])
    flight = ([SimpleBalloonState(*[float(v) for v in path_h[t, :, e]]) for t in range(steps)]
              if path_h is not None else [])
    results.append(EvaluationResult(
        seed=int(seed), cumulative_reward=float(res['cumulative_reward'][e]),
        time_within_radius=float(res['time_within_radius'][e]), out_of_power=bool(res['out_of_power'][e]),
        envelope_burst=bool(res['envelope_burst'][e]), zeropressure=bool(res['zeropressure'][e]),
        final_timestep=steps, flight_path=flight))
  return results


def results_to_json(results: Sequence[EvaluationResult]) -> str:
  """What eval/eval.py:87-96 writes."""
  return json.dumps(list(results), cls=EvalResultEncoder)


def combine_shards(shards: Sequence[str]) -> str:
  """eval/combine_eval_shards.py:38-59: concatenates per-shard JSON lists, ordered by seed."""
  merged = []
  for text in shards:
    merged.extend(json.loads(text))
  merged.sort(key=lambda r: r['seed'])
  return json.dumps(merged)
