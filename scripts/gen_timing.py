"""Times ble_generate_fields (the reset path's field writer) and reports its write bandwidth.

    python scripts/gen_timing.py --fields 65536 [--layout x64]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env, models   # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--fields', type=int, default=65536)
  ap.add_argument('--layout', default='x64')
  ap.add_argument('--decoder-precision', default='tf32')
  ap.add_argument('--reps', type=int, default=3)
  a = ap.parse_args()
  n = a.fields
  arena = batched_env.BatchedBalloonArena(n, precision='fp32', field_layout=a.layout, decoder_precision=a.decoder_precision)
  arena.set_decoder(models.load_decoder(''))
  arena.alloc_wind_fields(n)
  seeds = torch.arange(n, dtype=torch.int64, device='cuda')
  arena.sample_wind_fields(seeds)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for r in range(a.reps):
    arena.sample_wind_fields(seeds + 1 + r)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / a.reps
  per_field = {'x64': 1935360, 'x128': 3686400}[a.layout]
  print(json.dumps({'fields': n, 'layout': a.layout, 'decoder_precision': a.decoder_precision, 'generate_fields_ms': ms,
                    'window_bytes': per_field * n, 'write_gbs': per_field * n / (ms * 1e-3) / 1e9,
                    'fields_per_s': n / (ms * 1e-3)}))
  arena.close()


if __name__ == '__main__':
  main()
