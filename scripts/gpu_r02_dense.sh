#!/bin/bash
O=gpurun_out/r02dense; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_learner.py -m gpu -q 2>&1 | tail -15 | tee $O/pytest_dense.log
for be in tcgen05 cublas; do
  timeout 600 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --dense-backend $be 2> $O/train_$be.err | tail -1 | tee $O/train_qrdqn_n1_$be.json
done
