#!/bin/bash
mkdir -p gpurun_out/r02x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_step or rollout or free_running or full_size or noise or host_step or fp32_rollout" > gpurun_out/r02x/pytest.log 2>&1; tail -3 gpurun_out/r02x/pytest.log
timeout 300 python scripts/step_timing.py --sizes 4096,8192,16384,32768,65536 --variants fusedauto --steps 400 2>&1 | tee gpurun_out/r02x/timing.jsonl
