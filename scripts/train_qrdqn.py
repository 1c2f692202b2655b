"""QR-DQN training on vectorised balloons (train_acme_qrdqn.py:43-81; BASELINE configs[4]).

    python scripts/train_qrdqn.py --num-envs 4096 --iterations 40
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train_qrdqn.py --num-envs 32768    # 4,096 per GPU

Every rank flies `num-envs / world` balloons with the Perciatelli observation computed on the device
(decoder-generated wind field per episode), owns a replay ring in its own HBM and a learner replica; the
gradient is summed with one NCCL all-reduce per learner step.  The reference runs one 32-sample SGD step
per 4 environment steps of ONE balloon; with N balloons per iteration the same samples-per-insert ratio
(8) is kept by `--learner-steps` x `--batch-size` = 8 N per iteration unless overridden.
Prints one JSON line with the measured rates (CUDA events on the rank's stream, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from balloon_learning_environment_b200 import BatchedBalloonEnv, learner as learner_lib, models, sharding   # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--num-envs', type=int, default=4096, help='balloons in total (split over the ranks)')
  ap.add_argument('--iterations', type=int, default=40, help='timed iterations (one = every balloon steps once)')
  ap.add_argument('--warmup', type=int, default=8)
  ap.add_argument('--batch-size', type=int, default=0, help='per rank and learner step (0: 8 N / learner-steps)')
  ap.add_argument('--learner-steps', type=int, default=4, help='SGD steps per iteration')
  ap.add_argument('--replay-steps', type=int, default=0, help='ring length in steps (0: 2,000,000 / N, at least 16)')
  ap.add_argument('--decoder', default='', help='offlineskies22_decoder.msgpack (default: random-init weights)')
  ap.add_argument('--max-episode-length', type=int, default=960)
  ap.add_argument('--seed', type=int, default=0)
  ap.add_argument('--fp32-matmul', action='store_true', help='dense layers in fp32 FMA instead of TF32 tensor cores')
  ap.add_argument('--dense-backend', default='tcgen05', choices=('tcgen05', 'cublas'),
                  help="tcgen05: the hand-written ble_dense_tf32 kernel for forward and backward; cublas: nn.Linear + autograd (A/B)")
  args = ap.parse_args()
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  device = torch.device(f'cuda:{local}')
  torch.cuda.set_device(device)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)
  begin, end = sharding.shard_range(args.num_envs, rank, world)
  n = end - begin
  learner_steps = max(1, args.learner_steps)
  batch = args.batch_size or max(32, 8 * n // learner_steps)
  cfg = learner_lib.QrDqnConfig(batch_size=batch, max_episode_length=args.max_episode_length, min_replay_size=8 * n,
                                tf32_matmul=not args.fp32_matmul, dense_backend=args.dense_backend)

  layout = 'x128' if n * 3686400 <= 60e9 else 'x64'
  env = BatchedBalloonEnv(n, device=str(device), observation='perciatelli', field_layout=layout, seed=args.seed + rank,
                          decoder_params=models.load_decoder(args.decoder) if rank == 0 else None)   # rank 0 reads, NCCL broadcast
  learner = learner_lib.QrDqnLearner(cfg, device=device, seed=args.seed)          # same seed: identical replicas
  explore = learner_lib.MarcoPoloExploration(n, exploratory_episode_probability=cfg.exploratory_episode_probability,
                                             seed=args.seed + 17 * rank, device=device)
  ring = args.replay_steps or max(16, cfg.max_replay_size // args.num_envs)
  replay = learner_lib.DeviceReplay(n, ring, n_step=cfg.n_step, gamma=cfg.discount, device=device, seed=args.seed + rank)

  launches0 = env.arena.launch_count
  loop = learner_lib.TrainingLoop(env, learner, replay=replay, exploration=explore,
                                  learner_steps_per_iteration=learner_steps, seed=args.seed + rank)
  loop.run(args.warmup)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  stats = loop.run(args.iterations)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  red = sharding.reduce_run_stats(ms, n * args.iterations, env.arena.launch_count - launches0, device=device)
  checksum = learner.flat.double().sum()
  if world > 1:
    lo, hi = checksum.clone(), checksum.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    replicas_equal = bool(lo == hi)
  else:
    replicas_equal = True
  loop.run(4, profile=True)                        # untimed: where an iteration's time goes (rank 0's view)
  phases = {k: v / 4 for k, v in loop.phase_ms.items()}
  if rank == 0:
    sgd = stats['learner_steps']
    print(json.dumps({
        'workload': f'QR-DQN (8 x 600, 3 x 51 atoms), {args.num_envs} balloons with the Perciatelli observation, '
                    f'MarcoPolo exploration, device replay ring of {ring} steps',
        'n_gpus': world, 'envs_per_gpu': n, 'iterations': args.iterations, 'ms_per_iteration': red['elapsed_ms'] / args.iterations,
        'env_steps_per_s': red['env_steps'] / (red['elapsed_ms'] * 1e-3),
        'learner_steps_per_iteration': learner_steps, 'batch_size_per_gpu': batch,
        'dense_backend': args.dense_backend if not args.fp32_matmul else 'cublas',
        'learner_samples_per_s': sgd * batch * world / (red['elapsed_ms'] * 1e-3),
        'samples_per_insert': sgd * batch / (n * args.iterations),
        'last_loss': stats['last_loss'], 'mean_reward': stats['mean_reward'], 'episodes_rank0': stats['episodes'],
        'replicas_equal': replicas_equal, 'phase_ms_per_iteration': phases}), flush=True)
  env.close()
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
