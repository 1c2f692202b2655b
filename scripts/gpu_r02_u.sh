#!/bin/bash
mkdir -p gpurun_out/r02u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "wind_gather_matches or full_size" > gpurun_out/r02u/pytest.log 2>&1; grep -E "passed|failed|Error|error|wind gather|assert" gpurun_out/r02u/pytest.log | tail -12
