#!/bin/bash
# Round-1 re-entry verification: GPU parity tests, N=1 bench, L2 fetch-granularity experiment, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
rm -f gpurun_out/fetch_gran.jsonl
for g in 32 64 128; do
  echo "{\"fetch_granularity\": $g}" >> gpurun_out/fetch_gran.jsonl
  BLE_L2_FETCH_GRANULARITY=$g timeout 300 python scripts/gather_sweep.py --layouts x64 --fields 16384 >> gpurun_out/fetch_gran.jsonl 2>&1
done
cat gpurun_out/fetch_gran.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01b.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 2 > gpurun_out/bench_under_ncu.log 2>&1
tail -30 gpurun_out/launches_r01b.csv | cut -c1-200
