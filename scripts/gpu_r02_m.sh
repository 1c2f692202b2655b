#!/bin/bash
mkdir -p gpurun_out/r02m
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decoder or fused_field or latents" > gpurun_out/r02m/pytest_gen.log 2>&1; tail -3 gpurun_out/r02m/pytest_gen.log
for l in x64 x128; do timeout 300 python scripts/gen_timing.py --fields 32768 --layout $l | tee -a gpurun_out/r02m/gen_timing.jsonl; done
timeout 300 python scripts/gen_timing.py --fields 65536 | tee -a gpurun_out/r02m/gen_timing.jsonl
timeout 300 python scripts/gen_timing.py --fields 65536 --decoder-precision fp32 | tee -a gpurun_out/r02m/gen_timing.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_flow_to_windows -s 3 -c 1 -o gpurun_out/r02m/flow python scripts/gen_timing.py --fields 8192 --reps 1 > gpurun_out/r02m/ncu.log 2>&1
ncu -i gpurun_out/r02m/flow.ncu-rep --page details > gpurun_out/r02m/flow_details.txt 2>/dev/null
ncu -i gpurun_out/r02m/flow.ncu-rep --page source --csv > gpurun_out/r02m/flow_source.csv 2>/dev/null
rm -f gpurun_out/r02m/flow.ncu-rep
grep -E "Duration|Executed Ipc|Issue Slots Busy|DRAM Throughput|Memory Throughput|Warp Cycles Per Issued|Achieved Occupancy|Theoretical Occ" gpurun_out/r02m/flow_details.txt
