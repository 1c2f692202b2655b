#!/bin/bash
mkdir -p gpurun_out/r02s
for c in default 100 75 60 50 30; do
  if [ $c = default ]; then unset BLE_STEP_CARVEOUT; else export BLE_STEP_CARVEOUT=$c; fi
  echo "== carveout $c" | tee -a gpurun_out/r02s/timing.jsonl
  timeout 300 python scripts/step_timing.py --sizes 8192,32768,65536 --variants fusedauto --steps 400 2>&1 | tee -a gpurun_out/r02s/timing.jsonl
done
