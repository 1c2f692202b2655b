#!/bin/bash
mkdir -p gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_adaptor_under_reference.py -m gpu -q -x -k "features or incremental or column_tile or cuda_balloon_arena or eval or device" > gpurun_out/r02w/pytest_feat.log 2>&1; tail -3 gpurun_out/r02w/pytest_feat.log
V=$PWD/balloon_learning_environment_b200/variants
for rep in 1 2; do
for cfg in rr prev tile1 tile0; do
  unset BLE_B200_LIB BLE_COLUMN_TILE
  case $cfg in rr) export BLE_B200_LIB=$V/libble_rr.so;; prev) export BLE_B200_LIB=$V/libble_prev.so;; tile1) export BLE_COLUMN_TILE=1;; tile0) export BLE_COLUMN_TILE=0;; esac
  echo -n "$cfg " | tee -a gpurun_out/r02w/feature_timing_tile.jsonl
  timeout 300 python scripts/feature_timing.py --num-envs 65536 2>&1 | tail -1 | tee -a gpurun_out/r02w/feature_timing_tile.jsonl
done; done
