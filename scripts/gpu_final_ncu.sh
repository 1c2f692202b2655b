#!/bin/bash
# ncu --set full of the final build: the gather launch bench.py's roofline is quoted on, and the one-wave step
# kernels (k_noise + k_step_ws at 8,192 balloons = the per-GPU share of the 8-GPU strong-scaling split).
mkdir -p gpurun_out/final
timeout 240 ncu --set full --clock-control none --import-source on -k regex:k_wind_gather -s 4 -c 1 -o gpurun_out/final/gather \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 0 > gpurun_out/final/gather_bench_under_ncu.log 2>&1
ncu -i gpurun_out/final/gather.ncu-rep --page details > gpurun_out/final/gather_details.txt 2>/dev/null
ncu -i gpurun_out/final/gather.ncu-rep --page raw --csv > gpurun_out/final/gather_raw.csv 2>/dev/null
rm -f gpurun_out/final/gather.ncu-rep
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_step_ws|k_noise" -s 6 -c 2 -o gpurun_out/final/ws \
  python bench.py --num-envs 8192 --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 0 > gpurun_out/final/ws_bench_under_ncu.log 2>&1
ncu -i gpurun_out/final/ws.ncu-rep --page details > gpurun_out/final/ws_details.txt 2>/dev/null
rm -f gpurun_out/final/ws.ncu-rep
grep -E "dram__bytes_(read|write).sum\b|gpu__time_duration.sum" gpurun_out/final/gather_details.txt | head
grep -E "Duration|Registers Per|Executed Ipc Active|Achieved Occupancy" gpurun_out/final/ws_details.txt | head -12
ls -la gpurun_out/final
