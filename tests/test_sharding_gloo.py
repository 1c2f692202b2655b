"""world_size-2 gloo test of the multi-GPU host logic (runs on CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from balloon_learning_environment_b200 import sharding


def test_shard_range_partitions_exactly():
  for n in (1, 7, 65536, 262144, 65537):
    for g in (1, 2, 3, 4, 8):
      ranges = [sharding.shard_range(n, r, g) for r in range(g)]
      assert ranges[0][0] == 0 and ranges[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
      sizes = [e - b for b, e in ranges]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  b, e = sharding.shard_range(65536, rank, world)
  stats = sharding.reduce_run_stats(elapsed_ms=10.0 + 5.0 * rank, env_steps=(e - b) * 3, launches=6)
  pool = torch.full((2, 3), float(rank + 1))
  sharding.broadcast_field_pool(pool, src=0)
  # decoder weights: only rank 0 "read the checkpoint"
  params = None
  if rank == 0:
    g = torch.Generator().manual_seed(5)
    params = {f'Dense_{l}': {'kernel': torch.randn(i, o, generator=g).numpy(), 'bias': torch.randn(o, generator=g).numpy()}
              for l, (i, o) in enumerate(sharding.DECODER_SHAPES)}
  got = sharding.broadcast_decoder_params(params, 'cpu', src=0)
  checksum = sum(float(got[f'Dense_{l}']['kernel'].double().sum() + 3 * got[f'Dense_{l}']['bias'].double().sum())
                 for l in range(4))
  shapes = [tuple(got[f'Dense_{l}']['kernel'].shape) for l in range(4)]
  none_everywhere = sharding.broadcast_decoder_params(None, 'cpu', src=0) is None
  out[rank] = (stats, pool.clone(), (b, e), checksum, shapes, none_everywhere)
  dist.destroy_process_group()


def test_two_rank_reduction_and_broadcast():
  world = 2
  mgr = mp.Manager()
  out = mgr.dict()
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
  for rank in range(world):
    stats, pool, (b, e), checksum, shapes, none_everywhere = out[rank]
    assert checksum == out[0][3] and shapes == list(sharding.DECODER_SHAPES) and none_everywhere
    assert stats == {'elapsed_ms': 15.0, 'env_steps': 65536 * 3, 'launches': 12}     # max time, summed work
    assert torch.equal(pool, torch.ones(2, 3))                                      # rank 0's pool everywhere
    assert e - b == 32768
