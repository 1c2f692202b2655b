#!/bin/bash
# full GPU suite + smoke + the default bench line
mkdir -p gpurun_out/r02k
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k/pytest_gpu.log 2>&1; tail -8 gpurun_out/r02k/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02k/bench_n1.json 2> gpurun_out/r02k/bench_n1.err; tail -c 3000 gpurun_out/r02k/bench_n1.json
