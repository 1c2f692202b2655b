#!/bin/bash
# Last verification of the round: GPU parity tests (both step kernels against the reference goldens, the
# 65,536-balloon launch against the oracle), then the default bench line.
mkdir -p gpurun_out/final2
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/final2/pytest_gpu.log 2>&1; tail -15 gpurun_out/final2/pytest_gpu.log
timeout 100 python bench.py > gpurun_out/final2/bench_n1.json 2> gpurun_out/final2/bench_n1.err; tail -c 200 gpurun_out/final2/bench_n1.json
