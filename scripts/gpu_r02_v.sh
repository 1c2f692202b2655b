#!/bin/bash
mkdir -p gpurun_out/r02v
V=balloon_learning_environment_b200/variants
for lib in default mb12 mb8; do
  if [ $lib = default ]; then unset BLE_B200_LIB; else export BLE_B200_LIB=$PWD/$V/libble_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/r02v/timing.jsonl
  BLE_STEP_WARPS=0 timeout 300 python scripts/step_timing.py --sizes 16384,32768,65536 --variants fused0 --steps 400 2>&1 | tee -a gpurun_out/r02v/timing.jsonl
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "features or incremental" 2>&1 | tail -2
