#!/bin/bash
# 8-GPU evidence: BASELINE configs[2] (weak + strong split), configs[3] (262,144 balloons with the observation), configs[4] (QR-DQN).
N=${1:-8}
mkdir -p gpurun_out/multi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/multi/bench_n${N}.json 2> gpurun_out/multi/bench_n${N}.err
tail -c 400 gpurun_out/multi/bench_n${N}.json
timeout 400 $TR --master-port 29522 bench.py --gpus $N --num-envs 32768 --observation perciatelli --steps 20 --warmup 5 --scaling weak \
  > gpurun_out/multi/bench_obs_n${N}.json 2> gpurun_out/multi/bench_obs_n${N}.err
tail -c 400 gpurun_out/multi/bench_obs_n${N}.json
timeout 400 $TR --master-port 29523 scripts/train_qrdqn.py --num-envs 32768 --iterations 30 --warmup 130 > gpurun_out/multi/train_n${N}.json 2> gpurun_out/multi/train_n${N}.err
tail -1 gpurun_out/multi/train_n${N}.json
