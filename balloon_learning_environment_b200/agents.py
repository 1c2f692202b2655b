"""Batched baseline controllers behind the reference's Agent interface (agents/agent.py:38-162).

Every method takes / returns one entry PER BALLOON: observations float32 [N, 1099] (device), actions
int32 [N] (0 DOWN, 1 STAY, 2 UP).  The action rules run in CUDA (csrc/ble_agents.cuh).
"""
import enum
from typing import Optional, Sequence

import torch

from balloon_learning_environment_b200 import batched_env


class AgentMode(enum.Enum):            # agents/agent.py:29-35
  TRAIN = 'train'
  EVAL = 'eval'


class BatchedAgent:
  """agents/agent.py:38-149 for N simultaneous episodes."""

  def __init__(self, num_actions: int, observation_shape: Sequence[int], arena: batched_env.BatchedBalloonArena):
    self._num_actions = num_actions
    self._observation_shape = tuple(observation_shape)
    self._arena = arena
    self._mode = AgentMode.TRAIN

  def get_name(self) -> str:
    return self.__class__.__name__

  def begin_episode(self, observation: torch.Tensor) -> torch.Tensor:
    raise NotImplementedError

  def step(self, reward: torch.Tensor, observation: torch.Tensor) -> torch.Tensor:
    raise NotImplementedError

  def end_episode(self, reward: torch.Tensor, terminal: torch.Tensor) -> None:
    pass

  def set_mode(self, mode) -> None:
    self._mode = AgentMode(mode)

  def save_checkpoint(self, checkpoint_dir: str, iteration_number: int) -> None:
    pass

  def load_checkpoint(self, checkpoint_dir: str, iteration_number: int) -> int:
    return -1

  def reload_latest_checkpoint(self, checkpoint_dir: str) -> int:
    return -1


class RandomAgent(BatchedAgent):
  """agents/agent.py:152-162: uniform over the three commands."""

  def __init__(self, num_actions, observation_shape, arena, seed: Optional[int] = None):
    super().__init__(num_actions, observation_shape, arena)
    self._generator = torch.Generator(device=arena.device)
    if seed is not None:
      self._generator.manual_seed(int(seed))

  def _act(self):
    return torch.randint(0, self._num_actions, (self._arena.num_envs,), dtype=torch.int32,
                         device=self._arena.device, generator=self._generator)

  def begin_episode(self, observation):
    return self._act()

  def step(self, reward, observation):
    return self._act()


class StationSeekerAgent(BatchedAgent):
  """agents/station_seeker_agent.py:37-178; the score of all 361 levels and the argmax run in
  k_agent_station_seeker (one warp per balloon)."""

  def begin_episode(self, observation):
    return self._arena.station_seeker_actions(observation)

  def step(self, reward, observation):
    return self._arena.station_seeker_actions(observation)


class RandomWalkAgent(BatchedAgent):
  """agents/random_walk_agent.py:35-94: a Gaussian random walk of the target pressure per balloon."""

  def __init__(self, num_actions, observation_shape, arena, seed: Optional[int] = None):
    super().__init__(num_actions, observation_shape, arena)
    g = torch.Generator(device='cpu')
    if seed is not None:
      g.manual_seed(int(seed))
    self._seeds = torch.randint(0, 2**62, (arena.num_envs,), dtype=torch.int64, generator=g).to(arena.device)
    self._k = 0

  def begin_episode(self, observation):
    self._k = 0
    return self._arena.random_walk_actions(observation, self._seeds, 0)

  def step(self, reward, observation):
    self._k += 1
    return self._arena.random_walk_actions(observation, self._seeds, self._k)


REGISTRY = {'random': RandomAgent, 'station_seeker': StationSeekerAgent, 'random_walk': RandomWalkAgent}


def create_agent(name: str, num_actions: int, observation_shape, arena) -> BatchedAgent:
  """agents/agent_registry.py:40-75 for the controllers that exist here."""
  if name == 'quantile':                 # the QR-DQN agent brings the learner kernels with it: imported on demand
    from balloon_learning_environment_b200 import learner
    return learner.QuantileAgent(num_actions, observation_shape, arena)
  if name not in REGISTRY:
    raise ValueError(f'Unknown agent {name}; available: {sorted(REGISTRY)}')
  return REGISTRY[name](num_actions, observation_shape, arena)
