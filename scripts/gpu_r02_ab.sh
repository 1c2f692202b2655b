#!/bin/bash
mkdir -p gpurun_out/r02ab
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_step or rollout or free_running or full_size or noise or host_step or fp32_rollout or eval or features" > gpurun_out/r02ab/pytest.log 2>&1; tail -3 gpurun_out/r02ab/pytest.log
for pdl in 1 0; do
  echo "== BLE_STEP_PDL=$pdl" | tee -a gpurun_out/r02ab/timing.jsonl
  BLE_STEP_PDL=$pdl timeout 300 python scripts/step_timing.py --sizes 4096,8192,16384,65536 --variants fusedauto --steps 600 2>&1 | tee -a gpurun_out/r02ab/timing.jsonl
done
