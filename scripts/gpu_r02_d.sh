#!/bin/bash
# Round 2, call D: k_step_warp (throughput shape) + per-role loops in k_step_roles; parity, timings, ncu.
mkdir -p gpurun_out/r02d
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_step or rollout or free_running or host_step or full_size" > gpurun_out/r02d/pytest_sel.log 2>&1; tail -15 gpurun_out/r02d/pytest_sel.log
timeout 300 python scripts/step_timing.py --sizes 4096,8192,16384,65536 --variants fused0,fused4,fused8,fused14,thread --steps 300 2>&1 | tee gpurun_out/r02d/timing.jsonl
for spec in "65536 fused0" "8192 fused8"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_step_warp|k_step_roles" -s 8 -c 1 -o gpurun_out/r02d/k_$1 \
    python scripts/step_timing.py --sizes $1 --variants $2 --steps 10 > gpurun_out/r02d/ncu_$1.log 2>&1
  ncu -i gpurun_out/r02d/k_$1.ncu-rep --page details > gpurun_out/r02d/k_$1_details.txt 2>/dev/null
  ncu -i gpurun_out/r02d/k_$1.ncu-rep --page source --csv > gpurun_out/r02d/k_$1_source.csv 2>/dev/null
  rm -f gpurun_out/r02d/k_$1.ncu-rep
done
