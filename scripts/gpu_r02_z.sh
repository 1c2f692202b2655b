#!/bin/bash
mkdir -p gpurun_out/r02z
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_adaptor_under_reference.py -m gpu -q -x -s -k "features or incremental or column_tile or cuda_balloon_arena or eval or device" > gpurun_out/r02z/pytest_feat.log 2>&1; grep -E "passed|failed|worst|Error" gpurun_out/r02z/pytest_feat.log | tail -8
for rep in 1 2; do timeout 300 python scripts/feature_timing.py --num-envs 65536 2>&1 | tail -1 | tee -a gpurun_out/r02z/feature_timing.jsonl; done
