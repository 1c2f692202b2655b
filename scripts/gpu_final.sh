#!/bin/bash
# Final verification of the round's head on one B200: parity tests, smoke, both bench arms.
mkdir -p gpurun_out/final
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/final/gpu.txt
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; tail -2 gpurun_out/final/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 120 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err; tail -c 300 gpurun_out/final/bench_reference.json
timeout 300 python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err; tail -c 300 gpurun_out/final/bench_n1.json
