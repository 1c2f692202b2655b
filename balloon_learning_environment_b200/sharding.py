"""Multi-GPU layout of the batched environment: balloons are sharded, nothing else is shared.

Every balloon's transition reads only its own state, its own (or a read-only) wind field and its
own noise tables (`BalloonArena.step`, env/balloon_arena.py:184-202, has no inter-balloon term), so
rank r of G owns the contiguous range [r*N/G, (r+1)*N/G) and the step needs NO collective.
`torch.distributed` is used only to agree on timings / counters at the end of a run and, at
construction / reset, to hand every rank what rank 0 loaded from disk: the VAE decoder weights
(`broadcast_decoder_params`, called by BatchedBalloonEnv) and a shared wind-field pool
(`broadcast_field_pool`).  Over NCCL these are one NVSwitch broadcast of 22 MB (weights) or
3.7 MB per field; they are not on the step path.
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

DECODER_SHAPES = ((64, 1000), (1000, 1000), (1000, 1000), (1000, 4410))   # vae.Decoder (generative/vae.py:83-107)


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


def distributed_world() -> int:
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_decoder_params(params: Optional[dict], device, src: int = 0) -> Optional[dict]:
  """Every rank gets rank `src`'s decoder weights ({'Dense_i': {'kernel', 'bias'}}, the flax tree of
  offlineskies22_decoder.msgpack) or None if `src` has none.  Ranks other than `src` may pass None: only one
  process has to read the checkpoint.  A collective: every rank of the group must call it."""
  if distributed_world() == 1:
    return params
  have = torch.tensor([0 if params is None else 1], dtype=torch.int32, device=device)
  dist.broadcast(have, src=src)
  if int(have[0]) == 0:
    return None
  sizes = [i * o + o for i, o in DECODER_SHAPES]
  flat = torch.empty(sum(sizes), dtype=torch.float32, device=device)
  if dist.get_rank() == src:
    parts = []
    for l, (i, o) in enumerate(DECODER_SHAPES):
      layer = params[f'Dense_{l}']
      k = torch.as_tensor(np.asarray(layer['kernel'], np.float32)).reshape(-1)
      b = torch.as_tensor(np.asarray(layer['bias'], np.float32)).reshape(-1)
      if k.numel() != i * o or b.numel() != o:
        raise ValueError('decoder weights do not have the vae.Decoder shapes')
      parts += [k, b]
    flat.copy_(torch.cat(parts))
  dist.broadcast(flat, src=src)
  out, at = {}, 0
  for l, (i, o) in enumerate(DECODER_SHAPES):
    out[f'Dense_{l}'] = {'kernel': flat[at:at + i * o].view(i, o), 'bias': flat[at + i * o:at + i * o + o]}
    at += i * o + o
  return out
