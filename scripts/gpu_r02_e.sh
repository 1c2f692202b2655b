#!/bin/bash
# Round 2, call E: full GPU test-suite on the two-shape step kernels + bench line.
mkdir -p gpurun_out/r02e
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02e/pytest_gpu.log 2>&1; tail -15 gpurun_out/r02e/pytest_gpu.log
timeout 300 python scripts/step_timing.py --sizes 8192,65536 --variants fused0,fused4,fused8,fusedauto --steps 300 2>&1 | tee gpurun_out/r02e/timing.jsonl
timeout 900 python bench.py > gpurun_out/r02e/bench_n1.json 2> gpurun_out/r02e/bench_n1.err; tail -c 2500 gpurun_out/r02e/bench_n1.json; tail -5 gpurun_out/r02e/bench_n1.err
