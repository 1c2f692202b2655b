"""Evaluation suites (eval/suites.py:25-63).  The *_strata suites of the reference are seed lists
selected on ITS random streams (JAX threefry) and do not transfer; they are not provided."""
import dataclasses
from typing import List, Sequence


@dataclasses.dataclass
class EvaluationSuite:
  seeds: Sequence[int]
  max_episode_length: int


_SUITES = {
    'big_eval': (10_000, 960), 'medium_eval': (1_000, 960), 'small_eval': (100, 960),
    'tiny_eval': (10, 960), 'micro_eval': (1, 960),
}


def available_suites() -> List[str]:
  return list(_SUITES)


def get_eval_suite(name: str) -> EvaluationSuite:
  if name not in _SUITES:
    raise ValueError(f'Unknown eval suite {name}')
  count, length = _SUITES[name]
  return EvaluationSuite(list(range(count)), length)


def shard(suite: EvaluationSuite, shard_idx: int, num_shards: int) -> EvaluationSuite:
  """eval/eval.py:121-124."""
  start = int(len(suite.seeds) * shard_idx / num_shards)
  end = int(len(suite.seeds) * (shard_idx + 1) / num_shards)
  return EvaluationSuite(list(suite.seeds[start:end]), suite.max_episode_length)
