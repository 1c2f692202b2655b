#!/bin/bash
# Round 2, call C: noise v2 + smaller code; timings, ncu of k_step_fused at 65,536, free-running drift report.
mkdir -p gpurun_out/r02c
timeout 200 python scripts/step_timing.py --sizes 8192,65536 --variants fused4,fused8,thread --steps 300 2>&1 | tee gpurun_out/r02c/timing.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "free_running or single_step or noise or rollout" > gpurun_out/r02c/pytest_sel.log 2>&1; grep -v "^$" gpurun_out/r02c/pytest_sel.log | tail -70
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 8 -c 1 -o gpurun_out/r02c/fused_65536 \
  python scripts/step_timing.py --sizes 65536 --variants fused0 --steps 10 > gpurun_out/r02c/ncu.log 2>&1
ncu -i gpurun_out/r02c/fused_65536.ncu-rep --page details > gpurun_out/r02c/fused_65536_details.txt 2>/dev/null
ncu -i gpurun_out/r02c/fused_65536.ncu-rep --page source --csv > gpurun_out/r02c/fused_65536_source.csv 2>/dev/null
rm -f gpurun_out/r02c/fused_65536.ncu-rep
