#!/usr/bin/env python
"""Benchmark of the batched BLE transition function (BASELINE.json: env-steps/s at batch 65,536).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--num-envs 65536]

One "step" = one BalloonEnv.step (wind lookup + safety layers + 18 physics sub-steps + reward)
for every balloon of the batch.  Default workload = BASELINE.json configs[2]: 65,536 balloons,
random agent, one wind field per balloon (synthetic fields: the VAE decoder is reset-time only).
N > 1 (torchrun): balloons are sharded across ranks with no data-path collective.  Default is weak
scaling (65,536 balloons PER GPU, the single-GPU workload replicated; `value` = all balloons of all
ranks / max-over-ranks time); the same run then also times the strong-scaling split of ONE 65,536
batch (65,536 / N per GPU) and reports it under "strong_scaling".  `--scaling strong` makes the
fixed-total batch the headline instead.

`--impl reference` times the CPU oracle port (oracle/, the restated reference algorithm) on the
host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'env_steps_per_s'
UNIT = 'env-steps/s'
GATHER_BYTES_PER_LOOKUP = 156        # 16 B query + 4 B field index + 128 B corners + 8 B result
# DRAM bytes per k_wind_gather launch from the `ncu --set full` capture of the same launch shape
# (dram__bytes_read.sum + dram__bytes_write.sum); key = (field layout, lookups per launch).
NCU_GATHER_TRAFFIC = {('x64', 1 << 24): 3.652e9}      # 3.516 GB read + 0.136 GB written


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    with open(path) as f:
      return float(json.load(f)['hbm_gbs']), 'measured'
  return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
  """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.rows, self._stop_evt = index, [], threading.Event()

  def run(self):
    q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    while not self._stop_evt.is_set():
      try:
        out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
          self.rows.append([c.strip() for c in out.split(',')])
      except Exception:  # pylint: disable=broad-except
        pass
      self._stop_evt.wait(0.2)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=3)
    if not self.rows:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith('active') for r in self.rows)]
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
            'reasons': reasons, 'samples': len(self.rows)}


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle port): the only place bench.py touches oracle/
# ---------------------------------------------------------------------------------------------
def _oracle_worker(args):
  n, steps, seed = args
  os.environ.setdefault('OMP_NUM_THREADS', '1')
  from oracle import atmosphere, balloon, env as oenv, stable_init, wind
  rng = np.random.default_rng(seed)
  alpha = rng.uniform(0, 1, n)
  atm = atmosphere.Atmosphere(alpha)
  radius = 200e3 * rng.beta(1.2, 2.0, n); theta = rng.uniform(0, 2 * np.pi, n)
  pmax, _ = atm.at_height(np.full(n, 15240.0))
  b = balloon.make_batch(n, center_lat=np.radians(rng.uniform(-10, 10, n)), center_lng=np.radians(rng.uniform(-175, 175, n)),
                         date_time=1293840000 + rng.integers(0, 126144000, n), x=radius * np.cos(theta),
                         y=radius * np.sin(theta), pressure=rng.uniform(6500, pmax), upwelling_infrared=315.0)
  stable_init.cold_start_to_stable_params(b, atm)
  n_fields = min(n, 64)                                     # per-balloon fields: the CPU does not care
  fields = (rng.standard_normal((n_fields, 21, 21, 10, 9, 2)) * np.array([5.4, 1.5])).astype(np.float32)
  noise = wind.SimplexWindNoise(rng.integers(0, 1634753849, (n, 2, 5)), rng.uniform(-1, 1, (n, 2, 5, 4)))
  e = oenv.OracleEnv(oenv.OracleArena(b, atm, fields=fields, field_idx=rng.integers(0, n_fields, n), noise=noise))
  e.step(rng.integers(0, 3, n))                              # warm-up (imports, allocations)
  t0 = time.perf_counter()
  for _ in range(steps):
    e.step(rng.integers(0, 3, n))
  return time.perf_counter() - t0


def cpu_oracle_throughput(n_per_worker, steps, workers):
  """env-steps/s of the oracle port on `workers` host processes (1 = scalar baseline)."""
  if workers <= 1:
    dt = _oracle_worker((n_per_worker, steps, 0))
    return n_per_worker * steps / dt
  import multiprocessing as mp
  with mp.get_context('spawn').Pool(workers) as pool:
    times = pool.map(_oracle_worker, [(n_per_worker, steps, s) for s in range(workers)])
  return workers * n_per_worker * steps / max(times)


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  n_per_worker = 2048
  # warm-up + K steps, each step a bounded sample (cores x 2,048 balloons) of the 65,536-balloon batch
  t0 = time.perf_counter()
  value = cpu_oracle_throughput(n_per_worker, max(1, args.steps), cores)
  wall = time.perf_counter() - t0
  sample = f'{cores} processes x {n_per_worker} balloons x {max(1, args.steps)} steps of the oracle port (NumPy fp64)'
  # same workload description as the GPU arm's line (the host has one set of cores whatever --gpus says)
  n_total = args.num_envs * max(1, args.gpus) if args.scaling == 'weak' else args.num_envs
  line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'impl': 'reference', 'n_gpus': args.gpus, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': 1e3 * n_total / value, 'higher_is_better': True,
      'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
      'config': {'workload': f'batch={n_total} balloons, random agent, one synthetic wind field per balloon '
                             '+ simplex noise, 18 sub-steps per step (BASELINE configs[2]); the CPU arm times a bounded '
                             'sample of it',
                 'num_envs': n_total, 'envs_per_gpu': n_total // max(1, args.gpus), 'observation': 'none'},
      'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
      'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'wall_s': wall,
  }
  print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def upload_synthetic_fields(torch, arena, n_fields, device, seed, chunk=2048):
  """n_fields grids [21,21,10,9,2] fp32 with VAE-like magnitudes (u ~ 5.4 N(0,1), v ~ 1.5 N(0,1);
  SURVEY App. B), generated and converted chunk by chunk so only one chunk is resident twice."""
  g = torch.Generator(device=device); g.manual_seed(seed)
  scale = torch.tensor([5.4, 1.5], device=device)
  arena.alloc_wind_fields(n_fields)
  for s in range(0, n_fields, chunk):
    e = min(n_fields, s + chunk)
    arena.write_wind_fields(torch.randn(e - s, 21, 21, 10, 9, 2, generator=g, device=device) * scale, s)


def run_b200(args):
  import torch
  import torch.distributed as dist
  from balloon_learning_environment_b200 import batched_env, sharding

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local_rank}'))
  torch.cuda.set_device(local_rank)
  device = torch.device(f'cuda:{local_rank}')
  weak = args.scaling == 'weak'
  n_total = args.num_envs * world if weak else args.num_envs
  begin, end = sharding.shard_range(n_total, rank, world)    # contiguous balloon range of this rank
  n = end - begin

  with_obs = args.observation == 'perciatelli'
  n_fields = n if args.shared_fields == 0 else args.shared_fields
  layout = args.field_layout
  if layout == 'auto':      # one 128-byte line per lookup while the bank stays inside the ~64 GB TLB reach
    layout = 'x128' if n_fields * 3686400 <= 60e9 else 'x64'
  arena = batched_env.BatchedBalloonArena(n, device=str(device), precision='fp32', wind_model='grid', enable_noise=True,
                                          field_layout=layout, enable_features=True)
  obs_buf = torch.empty(n, 1099, dtype=torch.float32, device=device)
  upload_synthetic_fields(torch, arena, n_fields, device, seed=1234 + rank)
  arena.set_field_map(torch.arange(n, dtype=torch.int32, device=device) % n_fields)
  torch.cuda.empty_cache()
  g = torch.Generator(device='cpu'); g.manual_seed(2024 + rank)
  arena.reset(torch.randint(0, 2**62, (n,), dtype=torch.int64, generator=g))

  arena.features_track(with_obs)         # the plain rollout does not read the observation: no measurement kernel
  total = args.warmup + args.steps
  gd = torch.Generator(device=device); gd.manual_seed(7 + rank)
  actions = torch.randint(0, 3, (total, n), dtype=torch.int32, device=device, generator=gd)   # RandomAgent
  actions_host = actions.cpu().numpy()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def one_step(t):
    arena.step(actions[t])
    if with_obs:
      arena.features(obs_buf)

  # ---- device-resident timing -------------------------------------------------------------
  prefill = 0
  if with_obs:               # fill the 6 h WindGP window (120 measurements) first: the timed steps then pay the
    prefill = args.observation_prefill      # steady-state observation cost (one drop + one append per step)
    for t in range(prefill):
      arena.step(actions[t % total])
    arena.features(obs_buf)
  for t in range(args.warmup):
    one_step(t)
  barrier()
  launches0 = arena.launch_count
  sampler = ClockSampler(local_rank); sampler.start()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ev0.record()
  for t in range(args.warmup, total):
    one_step(t)
  ev1.record()
  barrier()
  ms = ev0.elapsed_time(ev1)
  launches = arena.launch_count - launches0
  live_frac = float((arena.get_state_dict()['status'] == 0).float().mean())

  # ---- end to end through the host-buffer C ABI call ----------------------------------------
  reward_h = np.zeros(n, np.float32); done_h = np.zeros(n, np.uint8)
  obs_h = torch.empty(n, 1099, dtype=torch.float32).pin_memory() if with_obs else None
  e2e_steps = args.steps

  def one_step_host(t):
    arena.step_host(actions_host[t], reward_h, done_h)
    if with_obs:                                   # observation computed on the device, read back to the host
      arena.features(obs_buf)
      obs_h.copy_(obs_buf, non_blocking=True)
      torch.cuda.current_stream().synchronize()

  for t in range(min(3, args.warmup)):
    one_step_host(t)
  barrier()
  t0 = time.perf_counter()
  for t in range(e2e_steps):
    one_step_host(args.warmup + t % args.steps)
  torch.cuda.synchronize()
  e2e_s = time.perf_counter() - t0
  clocks = sampler.stop()          # sampled across both timed regions (device-resident and end-to-end)

  # ---- the same step WITH the Perciatelli observation (reference path A), a few steps ----------
  obs_ms = None
  if not with_obs and args.observation_probe > 0:
    arena.features_clear(); arena.features_track(True); arena.features_observe()
    for t in range(args.observation_prefill):                          # fill the 6 h WindGP window first
      arena.step(actions[t % total])
    for _ in range(2):
      arena.step(actions[0]); arena.features(obs_buf)
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for t in range(args.observation_probe):
      arena.step(actions[t % total]); arena.features(obs_buf)
    o1.record(); torch.cuda.synchronize()
    obs_ms = o0.elapsed_time(o1) / args.observation_probe
    obs_live = float((arena.get_state_dict()['status'] == 0).float().mean())

  # ---- wind-gather roofline (dominant HBM kernel named by BASELINE.json's metric) -----------
  m = max(n * 8, 1 << 24)                                   # >= 16.7 M lookups, 2.6 GB of algorithmic traffic
  gq = torch.Generator(device=device); gq.manual_seed(99 + rank)
  xyzt = torch.empty(m, 4, dtype=torch.float32, device=device)
  xyzt[:, 0].uniform_(-500, 500, generator=gq); xyzt[:, 1].uniform_(-500, 500, generator=gq)
  xyzt[:, 2].uniform_(5000, 14000, generator=gq); xyzt[:, 3].uniform_(0, 48, generator=gq)
  # lookups grouped by balloon/field (the order every caller of the path issues them in: a balloon
  # queries its own field); the random-field order is measured separately below
  per_field = max(1, m // n_fields)
  fidx = torch.clamp(torch.arange(m, device=device) // per_field, max=n_fields - 1).to(torch.int32)
  for _ in range(3):
    arena.wind_forecast(xyzt, fidx)
  torch.cuda.synchronize()
  g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 10
  g0.record()
  for _ in range(reps):
    arena.wind_forecast(xyzt, fidx)
  g1.record()
  torch.cuda.synchronize()
  gather_ms = g0.elapsed_time(g1) / reps
  gather_gbs = GATHER_BYTES_PER_LOOKUP * m / (gather_ms * 1e-3) / 1e9
  fidx_rand = (torch.randperm(m, device=device, generator=gq) % n_fields).to(torch.int32)
  for _ in range(2):
    arena.wind_forecast(xyzt, fidx_rand)
  torch.cuda.synchronize()
  g0.record()
  for _ in range(reps):
    arena.wind_forecast(xyzt, fidx_rand)
  g1.record()
  torch.cuda.synchronize()
  gather_rand_ms = g0.elapsed_time(g1) / reps
  gather_rand_gbs = GATHER_BYTES_PER_LOOKUP * m / (gather_rand_ms * 1e-3) / 1e9

  # ---- reset path: per-episode cost, not part of `value` -----------
  del xyzt, fidx, fidx_rand
  torch.cuda.empty_cache()
  r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  seeds2 = torch.randint(0, 2**62, (n,), dtype=torch.int64, generator=g)
  arena.reset(seeds2); torch.cuda.synchronize()
  r0.record(); arena.reset(seeds2); r1.record(); torch.cuda.synchronize()
  reset_ms = r0.elapsed_time(r1)
  gw = torch.Generator(device=device); gw.manual_seed(5 + rank)
  dims = [64, 1000, 1000, 1000, 4410]
  arena.set_decoder({f'Dense_{i}': {'kernel': (torch.randn(dims[i], dims[i + 1], generator=gw, device=device)
                                               * (2.0 / dims[i]) ** 0.5).cpu().numpy(),
                                    'bias': np.zeros(dims[i + 1], np.float32)} for i in range(4)})
  z = torch.randn(2048, 64, generator=gw, device=device)
  arena.decode_wind_fields(z); torch.cuda.synchronize()
  r0.record(); arena.decode_wind_fields(z); r1.record(); torch.cuda.synchronize()
  decode_ms = r0.elapsed_time(r1)
  del z
  # BASELINE configs[2] draws a new VAE field for every balloon at every reset: all n fields through
  # ble_generate_fields (latents -> 4 GEMMs -> resize / curl -> 128-byte windows in the bank)
  gen_seeds = torch.randint(0, 2**62, (n_fields,), dtype=torch.int64, generator=g)
  arena.sample_wind_fields(gen_seeds[:2048]); torch.cuda.synchronize()
  r0.record(); arena.sample_wind_fields(gen_seeds); r1.record(); torch.cuda.synchronize()
  generate_ms = r0.elapsed_time(r1)
  reset_path = {'reset_ms': reset_ms, 'balloons': n,
                'what': 'ble_reset: Philox sampling, stable init, sunrise/sunset search, 10 noise permutation tables per balloon',
                'generate_fields_ms': generate_ms, 'generated_fields': n_fields,
                'episode_steps': 960,
                'amortised_env_steps_per_s': n * 960 / ((960 * ms / args.steps + reset_ms + generate_ms) * 1e-3),
                'amortised_note': 'this rank: 960-step episodes with ble_generate_fields (one new field per balloon) + '
                                  'ble_reset charged once per episode',
                'decoder_fields_per_s': 2048 / (decode_ms * 1e-3), 'decoder_batch': 2048,
                'decoder': 'ble_decode_fields: 4 cuBLASLt fp32 GEMMs (64-1000-1000-1000-4410) + resize/curl epilogue, '
                           'random-init weights'}

  # ---- N > 1, weak run: also time the strong-scaling split of ONE --num-envs batch ----------------
  strong = None
  if world > 1 and weak:
    arena.close()
    del obs_buf
    torch.cuda.empty_cache()
    sb, se = sharding.shard_range(args.num_envs, rank, world)
    ns = se - sb
    s_layout = 'x128' if ns * 3686400 <= 60e9 else 'x64'
    arena = batched_env.BatchedBalloonArena(ns, device=str(device), precision='fp32', wind_model='grid',
                                            enable_noise=True, field_layout=s_layout)
    upload_synthetic_fields(torch, arena, ns, device, seed=4321 + rank)
    arena.set_field_map(torch.arange(ns, dtype=torch.int32, device=device))
    arena.reset(torch.randint(0, 2**62, (ns,), dtype=torch.int64, generator=g))
    s_actions = actions[:, :ns].contiguous()
    for t in range(args.warmup):
      arena.step(s_actions[t])
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(args.warmup, total):
      arena.step(s_actions[t])
    s1.record()
    barrier()
    s_ms = sharding.reduce_run_stats(s0.elapsed_time(s1), 0, 0, device=device)['elapsed_ms']
    strong = {'num_envs': args.num_envs, 'envs_per_gpu': ns, 'field_layout': s_layout, 'ms_per_step': s_ms / args.steps,
              'value': args.num_envs * args.steps / (s_ms * 1e-3), 'unit': UNIT,
              'note': 'one --num-envs batch split over the GPUs (fixed total work), device-resident, max over ranks'}

  stats = sharding.reduce_run_stats(ms, n * args.steps, launches, device=device)
  ms, launches = stats['elapsed_ms'], stats['launches']
  e2e_s = sharding.reduce_run_stats(e2e_s * 1e3, 0, 0, device=device)['elapsed_ms'] * 1e-3
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return

  peak, peak_src = load_peaks()
  value = n_total * args.steps / (ms * 1e-3)
  e2e_value = n_total * e2e_steps / e2e_s
  line = {
      'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
      'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': f'batch={n_total} balloons, random agent, one synthetic wind field per balloon '
                             f'({n_fields} fields/GPU) + simplex noise, 18 sub-steps per step '
                             + ('+ the Perciatelli observation every step (BASELINE configs[3])' if with_obs
                                else '(BASELINE configs[2])'),
                 'num_envs': n_total, 'envs_per_gpu': n, 'fields_per_gpu': n_fields, 'field_layout': layout,
                 'l2': 'inputs larger than L2 (per-balloon fields + 2.5 KB noise tables per balloon)',
                 'observation': args.observation, 'observation_prefill_steps': prefill,
                 'live_fraction_after_run': live_frac},
      'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': 4 * n_total,
              'd2h_bytes_per_step': (5 + (4396 if with_obs else 0)) * n_total,
              'api': 'ble_step_host (host int32 actions in, host float32 reward + uint8 done out; the step kernel reads / writes the pinned staging buffers over PCIe)'
                     + (' + ble_features_perciatelli read back to pinned host memory' if with_obs else '')},
      'gpu_launches': launches,
      'clocks': clocks,
      'roofline': {'kernel': 'k_wind_gather', 'bound': 'hbm', 'achieved': gather_gbs, 'peak': peak, 'unit': 'GB/s',
                   'frac': gather_gbs / peak, 'traffic': NCU_GATHER_TRAFFIC.get((layout, m)),
                   'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, '
                                   'profiles/r01c_k_wind_gather_details.txt)', 'peak_source': peak_src,
                   'lookups_per_launch': m, 'bytes_per_lookup': GATHER_BYTES_PER_LOOKUP, 'ms_per_launch': gather_ms,
                   'access': f'{m} uniformly random (x, y, p, t) points, grouped by field ({per_field} per field, '
                             f'{n_fields} fields, layout {layout})',
                   'random_field_order': {'achieved': gather_rand_gbs, 'frac': gather_rand_gbs / peak,
                                          'ms_per_launch': gather_rand_ms,
                                          'note': 'same lookups with the field chosen at random per lookup; above a '
                                                  '~64 GB field bank this is TLB-miss bound (DESIGN.md section 4)'}},
  }
  line['reset_path'] = reset_path
  if obs_ms is not None:
    line['with_perciatelli_observation'] = {'ms_per_step': obs_ms, 'value': n / (obs_ms * 1e-3), 'unit': UNIT,
                                            'steps': args.observation_probe,
                                            'live_fraction': obs_live,
                                            'note': 'ble_step + ble_features_perciatelli (1099 float32 features per balloon, '
                                                    'full 120-measurement WindGP window), device-resident; this rank\'s '
                                                    'balloons only'}
  if strong is not None:
    line['strong_scaling'] = strong
  if world == 1 and not args.no_cpu_baseline:
    t0 = time.perf_counter()
    v1 = cpu_oracle_throughput(2048, 40, 1)
    line['cpu_baseline'] = {'value': v1, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                            'sample': '1 process x 2048 balloons x 40 steps of the oracle port (NumPy fp64), '
                                      f'{time.perf_counter() - t0:.1f} s'}
  print(json.dumps(line), flush=True)
  arena.close()
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--num-envs', type=int, default=65536,
                  help='balloons per GPU (--scaling weak) or in total (--scaling strong)')
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                  help='N > 1: weak = --num-envs balloons per GPU; strong = --num-envs balloons split over the GPUs')
  ap.add_argument('--shared-fields', type=int, default=0, help='0 = one field per balloon; else size of a shared pool')
  ap.add_argument('--field-layout', default='auto', choices=['auto', 'x64', 'x128'])
  ap.add_argument('--observation', default='none', choices=['none', 'perciatelli'],
                  help="'perciatelli': every step also computes the 1099-feature observation")
  ap.add_argument('--observation-probe', type=int, default=10,
                  help='extra steps timed WITH the observation when --observation none (0 = skip)')
  ap.add_argument('--observation-prefill', type=int, default=125,
                  help='untimed steps that fill the 6 h WindGP window before the observation is timed')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
