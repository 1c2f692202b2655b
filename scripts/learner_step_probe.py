"""Times QrDqnLearner.step alone (batch 8192, the 8 x 600 network) per dense backend; under ncu (--once) it is the
launch list of ONE step.    python scripts/learner_step_probe.py [--backend tcgen05|cublas] [--once] [--eager]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import learner as lrn  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--backend', default='tcgen05')
  ap.add_argument('--batch', type=int, default=8192)
  ap.add_argument('--once', action='store_true')
  ap.add_argument('--eager', action='store_true')
  args = ap.parse_args()
  cfg = lrn.QrDqnConfig(dense_backend=args.backend, cuda_graph=not args.eager)
  learner = lrn.QrDqnLearner(cfg, seed=0)
  g = torch.Generator(device='cuda'); g.manual_seed(0)
  b = args.batch
  batch = {'state': torch.rand(b, 1099, device='cuda', generator=g), 'next_state': torch.rand(b, 1099, device='cuda', generator=g),
           'action': torch.randint(0, 3, (b,), dtype=torch.int32, device='cuda', generator=g),
           'return': torch.rand(b, device='cuda', generator=g), 'discount': torch.full((b,), 0.965, device='cuda'),
           'valid': torch.ones(b, dtype=torch.uint8, device='cuda')}
  if args.once:
    learner.step(batch); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    learner.step(batch); torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    return
  for _ in range(5):
    learner.step(batch)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  steps = 50
  e0.record()
  for _ in range(steps):
    learner.step(batch)
  e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / steps
  flop = 0
  dims = [1099] + [600] * 7 + [153]
  for i, o in zip(dims[:-1], dims[1:]):
    flop += 2 * b * i * o * 2          # online + target forward
    flop += 2 * b * i * o              # weight gradient
    flop += 2 * b * i * o if i != 1099 else 0   # input gradient
  print(json.dumps({'backend': args.backend, 'cuda_graph': not args.eager, 'batch': b, 'ms_per_sgd_step': ms,
                    'dense_tflops': flop / ms / 1e9}), flush=True)


if __name__ == '__main__':
  main()
