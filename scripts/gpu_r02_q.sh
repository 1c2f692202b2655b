#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 300 python scripts/step_timing.py --sizes 65536 --variants fused0 --steps 400 2>&1 | tee gpurun_out/r02q/timing.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_warp -s 20 -c 1 -o gpurun_out/r02q/k python scripts/step_timing.py --sizes 65536 --variants fused0 --steps 40 > gpurun_out/r02q/ncu.log 2>&1
ncu -i gpurun_out/r02q/k.ncu-rep --page details > gpurun_out/r02q/k_details.txt 2>/dev/null
ncu -i gpurun_out/r02q/k.ncu-rep --page source --csv > gpurun_out/r02q/k_source.csv 2>/dev/null
rm -f gpurun_out/r02q/k.ncu-rep
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Warp Cycles Per Issued|Executed Instructions|Achieved Occ|Local" gpurun_out/r02q/k_details.txt | head
