#!/bin/bash
mkdir -p gpurun_out/r02o
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "features or incremental or cuda_balloon_arena or eval" > gpurun_out/r02o/pytest_feat.log 2>&1; grep -E "passed|failed|worst|error" gpurun_out/r02o/pytest_feat.log | tail -8
timeout 300 python scripts/feature_timing.py --num-envs 65536 2>&1 | tee gpurun_out/r02o/feature_timing.json | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gp_posterior -s 2 -c 1 -o gpurun_out/r02o/gp python scripts/feature_timing.py --num-envs 16384 --fields 1024 > gpurun_out/r02o/ncu.log 2>&1
ncu -i gpurun_out/r02o/gp.ncu-rep --page details > gpurun_out/r02o/gp_details.txt 2>/dev/null
ncu -i gpurun_out/r02o/gp.ncu-rep --page source --csv > gpurun_out/r02o/gp_source.csv 2>/dev/null
rm -f gpurun_out/r02o/gp.ncu-rep
grep -E "Duration|Executed Ipc|Issue Slots Busy|Warp Cycles Per Issued|Executed Instructions" gpurun_out/r02o/gp_details.txt
