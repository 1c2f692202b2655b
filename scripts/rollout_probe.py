"""ble_rollout per-step time as a function of K at 65,536 balloons (why is K = 32 slower than 32 launches?)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
  dev = torch.device('cuda:0')
  a = batched_env.BatchedBalloonArena(n, precision='fp32', enable_noise=True, field_layout='x128')
  g = torch.Generator(device=dev); g.manual_seed(1)
  nf = 2048
  a.alloc_wind_fields(nf)
  for s in range(0, nf, 1024):
    a.write_wind_fields(torch.randn(1024, 21, 21, 10, 9, 2, generator=g, device=dev) * 5.0, s)
  a.set_field_map(torch.arange(n, dtype=torch.int32, device=dev) % nf)
  a.reset(torch.arange(n, dtype=torch.int64) * 7 + 1)
  actions = torch.randint(0, 3, (64, n), dtype=torch.int32, device=dev, generator=g)
  out = {}
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  for k in (1, 2, 4, 8, 16, 32, 64):
    a.rollout(actions[:k]); torch.cuda.synchronize()
    reps = max(2, 256 // k)
    e0.record()
    for r in range(reps):
      a.rollout(actions[:k])
    e1.record(); torch.cuda.synchronize()
    out[f'K{k}'] = e0.elapsed_time(e1) / (reps * k) * 1e3
  for t in range(5):
    a.step(actions[t])
  torch.cuda.synchronize()
  e0.record()
  for t in range(256):
    a.step(actions[t % 64])
  e1.record(); torch.cuda.synchronize()
  out['step'] = e0.elapsed_time(e1) / 256 * 1e3
  print(json.dumps({'n': n, 'us_per_step': out}))
  a.close()


if __name__ == '__main__':
  main()
