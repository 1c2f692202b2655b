import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from balloon_learning_environment_b200 import batched_env
dev = torch.device('cuda:0'); n = 16384
arena = batched_env.BatchedBalloonArena(n, precision='fp32', enable_noise=True, enable_features=True)
bench.upload_synthetic_fields(torch, arena, 1024, dev, seed=3)
arena.set_field_map(torch.arange(n, dtype=torch.int32, device=dev) % 1024)
arena.reset(torch.arange(n, dtype=torch.int64) + 17)
g = torch.Generator(device=dev); g.manual_seed(0)
obs = torch.zeros(n, 1099, dtype=torch.float32, device=dev)
E = 7548
hist = []
def snap():
  st = arena.get_state_dict()
  w = arena.wind_at_balloon()[E].cpu().numpy()
  x, y, p, t = float(st['x'][E]), float(st['y'][E]), float(st['pressure'][E]), int(st['time_elapsed'][E])
  f = arena.wind_forecast(torch.tensor([[x/1000, y/1000, p, t/3600.0]]), torch.tensor([E % 1024], dtype=torch.int32))[0].cpu().numpy()
  hist.append([x, y, p, t, w[0]-f[0], w[1]-f[1]])
snap()
for t in range(81):
  arena.step(torch.randint(0, 3, (n,), dtype=torch.int32, device=dev, generator=g)); snap()
arena.features(obs)
row = obs[E].cpu().numpy(); col = row[16:].reshape(361, 3)
bad = ~np.isfinite(col).all(1)
print('bad slots', np.nonzero(bad)[0][:5], int(bad.sum()), 'first rows', col[np.nonzero(bad)[0][:3]], 'good reachable rows', col[(col[:,0]!=0)&~bad][:3])
h = np.array(hist, np.float64)
print('history tail', h[-3:], 'count', len(h), 'min dt', np.diff(h[:,3]).min(), 'err max', np.abs(h[:,4:]).max(), 'finite', np.isfinite(h).all())
from oracle import features as F
import scipy.linalg
a = h[:, :4] / F.GP_LENGTH_SCALE
k = F.GP_SIGMA2 * np.exp(-np.sqrt(((a[:, None] - a[None]) ** 2).sum(-1))); k[np.diag_indices_from(k)] += F.GP_NOISE
print('cond', np.linalg.cond(k), 'min eig', np.linalg.eigvalsh(k).min())
