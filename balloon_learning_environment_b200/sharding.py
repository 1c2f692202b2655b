"""Multi-GPU layout of the batched environment: balloons are sharded, nothing else is shared.

Every balloon's transition reads only its own state, its own (or a read-only) wind field and its
own noise tables (`BalloonArena.step`, env/balloon_arena.py:184-202, has no inter-balloon term), so
rank r of G owns the contiguous range [r*N/G, (r+1)*N/G) and the step needs NO collective.
`torch.distributed` is used only to agree on timings / counters at the end of a run and, when
asked, to broadcast a shared wind-field pool at reset.
"""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous [begin, end) of balloons owned by `rank`; the first N % G ranks get one extra."""
  if not 0 <= rank < world_size:
    raise ValueError('rank out of range')
  base, extra = divmod(num_envs, world_size)
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def reduce_run_stats(elapsed_ms: float, env_steps: int, launches: int, device=None) -> Dict[str, float]:
  """max-over-ranks time, sum-over-ranks work: the whole-job numbers rank 0 reports."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return {'elapsed_ms': float(elapsed_ms), 'env_steps': int(env_steps), 'launches': int(launches)}
  t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
  c = torch.tensor([env_steps, launches], dtype=torch.int64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  dist.all_reduce(c, op=dist.ReduceOp.SUM)
  return {'elapsed_ms': float(t[0]), 'env_steps': int(c[0]), 'launches': int(c[1])}


def broadcast_field_pool(fields: torch.Tensor, src: int = 0) -> torch.Tensor:
  """Reset-time only: make every rank fly the same pool of wind fields (NCCL/gloo broadcast)."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.broadcast(fields, src=src)
  return fields
