"""Evaluates a baseline controller on a suite, every seed as one balloon of a GPU batch (eval/eval.py:99-132).

    python scripts/run_eval.py --agent station_seeker --suite small_eval --output_dir /tmp/ble/eval
    torchrun --nproc-per-node 8 scripts/run_eval.py --suite big_eval ...      # one shard per GPU (eval.py:121-124)

Wind fields come from the VAE decoder (`--decoder path/to/offlineskies22_decoder.msgpack`, the reference's
weights) or, without it, from random-init weights of the same architecture (there is no dataset here).
Writes <output_dir>/<agent>_<shard>.json in the reference's schema and prints one summary line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import BatchedBalloonEnv, agents, eval_lib, models, suites   # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--agent', default='station_seeker', choices=sorted(agents.REGISTRY))
  ap.add_argument('--suite', default='small_eval', choices=suites.available_suites())
  ap.add_argument('--max-episode-length', type=int, default=0, help='override the suite (0 = keep)')
  ap.add_argument('--decoder', default='')
  ap.add_argument('--output_dir', default='')
  ap.add_argument('--no-flight-path', action='store_true')
  args = ap.parse_args()
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  suite = suites.shard(suites.get_eval_suite(args.suite), rank, world)
  if args.max_episode_length:
    suite.max_episode_length = args.max_episode_length
  device = f'cuda:{local}'
  torch.cuda.set_device(local)
  if world > 1:                       # rank 0 reads the decoder checkpoint; the others receive it over NCCL
    torch.distributed.init_process_group('nccl', device_id=torch.device(device))
  n = len(suite.seeds)
  layout = 'x128' if n * 3686400 <= 60e9 else 'x64'
  env = BatchedBalloonEnv(n, device=device, observation='perciatelli',
                          decoder_params=models.load_decoder(args.decoder) if rank == 0 else None, field_layout=layout)
  agent = agents.create_agent(args.agent, env.action_space.n, env.observation_space.shape, env.arena)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  results = eval_lib.eval_agent(agent, env, suite, calculate_flight_path=not args.no_flight_path)
  torch.cuda.synchronize()
  seconds = time.perf_counter() - t0
  if args.output_dir:
    os.makedirs(args.output_dir, exist_ok=True)
    with open(os.path.join(args.output_dir, f'{args.agent}_{rank}.json'), 'w') as f:
      f.write(eval_lib.results_to_json(results))
  steps = sum(r.final_timestep for r in results)
  print(json.dumps({'agent': args.agent, 'suite': args.suite, 'shard': rank, 'num_shards': world, 'seeds': n,
                    'max_episode_length': suite.max_episode_length, 'seconds': seconds, 'env_steps': steps,
                    'env_steps_per_s': steps / seconds,
                    'mean_cumulative_reward': float(np.mean([r.cumulative_reward for r in results])),
                    'mean_time_within_radius': float(np.mean([r.time_within_radius for r in results])),
                    'terminated': int(sum(r.out_of_power or r.envelope_burst or r.zeropressure for r in results))}),
        flush=True)
  env.close()
  if world > 1:
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
  main()
