#!/bin/bash
# Round 2, call A: parity of the fused step kernel, first timings, ncu of k_step_fused at 65,536 and 8,192 balloons.
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/r02a/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02a/pytest_parity.log 2>&1; tail -15 gpurun_out/r02a/pytest_parity.log
timeout 300 python scripts/step_timing.py > gpurun_out/r02a/step_timing.jsonl 2> gpurun_out/r02a/step_timing.err; cat gpurun_out/r02a/step_timing.jsonl; tail -3 gpurun_out/r02a/step_timing.err
timeout 600 python bench.py > gpurun_out/r02a/bench_n1.json 2> gpurun_out/r02a/bench_n1.err; tail -c 1500 gpurun_out/r02a/bench_n1.json; tail -5 gpurun_out/r02a/bench_n1.err
for n in 65536 8192; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 8 -c 1 -o gpurun_out/r02a/fused_$n \
    python scripts/step_timing.py --sizes $n --variants fused0 --steps 10 > gpurun_out/r02a/ncu_$n.log 2>&1
  ncu -i gpurun_out/r02a/fused_$n.ncu-rep --page details > gpurun_out/r02a/fused_${n}_details.txt 2>/dev/null
  ncu -i gpurun_out/r02a/fused_$n.ncu-rep --page raw --csv > gpurun_out/r02a/fused_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02a/fused_$n.ncu-rep --page source --csv > gpurun_out/r02a/fused_${n}_source.csv 2>/dev/null
  rm -f gpurun_out/r02a/fused_$n.ncu-rep
done
ls -la gpurun_out/r02a
