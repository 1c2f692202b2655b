#!/bin/bash
mkdir -p gpurun_out/r02p
V=balloon_learning_environment_b200/variants
for lib in default w2 w1; do
  if [ $lib = default ]; then unset BLE_B200_LIB; else export BLE_B200_LIB=$PWD/$V/libble_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/r02p/timing.jsonl
  timeout 300 python scripts/step_timing.py --sizes 32768,65536 --variants fused0 --steps 400 2>&1 | tee -a gpurun_out/r02p/timing.jsonl
done
