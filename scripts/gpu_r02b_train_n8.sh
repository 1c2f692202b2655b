#!/bin/bash
# BASELINE configs[4] on 8 GPUs with the final dense path (one learner replica + 4,096 balloons per GPU, NCCL all-reduce per SGD step)
O=gpurun_out/r02bmulti; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29552 scripts/train_qrdqn.py --num-envs 32768 --iterations 30 2> $O/train_n8.err | tail -1 > $O/train_qrdqn_n8_tcgen05.json; cut -c1-900 $O/train_qrdqn_n8_tcgen05.json
