#!/bin/bash
mkdir -p gpurun_out/r02y
V=balloon_learning_environment_b200/variants
for lib in default evict default evict; do
  if [ $lib = default ]; then unset BLE_B200_LIB; else export BLE_B200_LIB=$PWD/$V/libble_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/r02y/timing.jsonl
  BLE_STEP_WARPS=0 timeout 300 python scripts/step_timing.py --sizes 32768,65536 --variants fused0 --steps 400 2>&1 | tee -a gpurun_out/r02y/timing.jsonl
done
