"""Host-side cost of one ble_step call (Python + ctypes + launch), measured with a batch too small to matter on the GPU."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import batched_env


def main():
  out = {}
  for n in (32, 4096, 8192):
    a = batched_env.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True)
    a.reset(torch.arange(n, dtype=torch.int64))
    acts = torch.randint(0, 3, (64, n), dtype=torch.int32, device='cuda')
    rows = [acts[i] for i in range(64)]
    for mode in ('index', 'presliced'):
      for _ in range(50):
        a.step(rows[0])
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      k = 2000
      if mode == 'index':
        for t in range(k):
          a.step(acts[t % 64])
      else:
        for t in range(k):
          a.step(rows[t & 63])
      t1 = time.perf_counter()
      torch.cuda.synchronize()
      t2 = time.perf_counter()
      out[f'n{n}_{mode}'] = {'issue_us_per_step': (t1 - t0) / k * 1e6, 'total_us_per_step': (t2 - t0) / k * 1e6}
    a.close()
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  main()
