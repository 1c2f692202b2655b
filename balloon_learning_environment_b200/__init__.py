"""B200-native batched transition function of the Balloon Learning Environment.

Public surface:
  BatchedBalloonArena / BatchedBalloonEnv  -- N balloons on one GPU (tensor API)
  CudaBalloonArena                         -- N = 1 adaptor behind the reference's
                                              BalloonArenaInterface (env/balloon_arena.py:42-120)
  build()                                  -- compile csrc/ -> libble_b200.so (nvcc, sm_100a)

The CUDA library is required: importing the env classes and creating an arena fails loudly when
libble_b200.so or a CUDA device is missing.  Nothing in this package imports `oracle/`.
"""
from balloon_learning_environment_b200._build import build  # noqa: F401
from balloon_learning_environment_b200 import _lib  # noqa: F401

__all__ = ['build', 'BatchedBalloonArena', 'BatchedBalloonEnv', 'CudaBalloonArena']


def __getattr__(name):
  if name in ('BatchedBalloonArena', 'BatchedBalloonEnv'):
    from balloon_learning_environment_b200 import batched_env
    return getattr(batched_env, name)
  if name == 'CudaBalloonArena':
    from balloon_learning_environment_b200 import arena
    return arena.CudaBalloonArena
  raise AttributeError(name)
