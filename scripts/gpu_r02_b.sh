#!/bin/bash
# Round 2, call B: A/B of resident CTAs per SM for k_step_fused<4> (launch bounds 4 / 5 / 6) and the rest of the parity suite.
mkdir -p gpurun_out/r02b
for v in default mb4 mb6; do
  if [ $v != default ]; then export BLE_B200_LIB=$PWD/balloon_learning_environment_b200/variants/libble_$v.so; fi
  echo "== $v"
  timeout 200 python scripts/step_timing.py --sizes 8192,65536 --variants fused4,fused8 --steps 300 2>&1 | tee -a gpurun_out/r02b/timing_$v.jsonl
done
unset BLE_B200_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02b/pytest_parity.log 2>&1; tail -40 gpurun_out/r02b/pytest_parity.log
