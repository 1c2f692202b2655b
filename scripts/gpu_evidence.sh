#!/bin/bash
# Collects the round's evidence on one B200: tests, bench line, launch list, ncu details of the main kernels.
mkdir -p gpurun_out/ev
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/ev/pytest_gpu.log 2>&1; tail -2 gpurun_out/ev/pytest_gpu.log
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/ev/gpu.txt
timeout 600 python bench.py > gpurun_out/ev/bench_n1.json 2> gpurun_out/ev/bench_n1.err; tail -c 300 gpurun_out/ev/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 > gpurun_out/ev/bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/ev/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ev/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 3 > gpurun_out/ev/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_noise|k_wind_gather" -s 6 -c 6 -o gpurun_out/ev/step \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 0 > /dev/null 2>&1
ncu -i gpurun_out/ev/step.ncu-rep --page details > gpurun_out/ev/step_details.txt 2>/dev/null
ncu -i gpurun_out/ev/step.ncu-rep --page raw --csv > gpurun_out/ev/step_raw.csv 2>/dev/null
rm -f gpurun_out/ev/step.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gp_column3|k_gp_update|k_feat_range" -s 15 -c 3 -o gpurun_out/ev/gp \
  python scripts/feature_timing.py --num-envs 16384 --fields 1024 > /dev/null 2>&1
ncu -i gpurun_out/ev/gp.ncu-rep --page details > gpurun_out/ev/gp_details.txt 2>/dev/null
ncu -i gpurun_out/ev/gp.ncu-rep --page raw --csv > gpurun_out/ev/gp_raw.csv 2>/dev/null
rm -f gpurun_out/ev/gp.ncu-rep
timeout 300 python scripts/feature_timing.py > gpurun_out/ev/feature_timing.json 2>&1; tail -1 gpurun_out/ev/feature_timing.json
timeout 300 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --warmup 130 > gpurun_out/ev/train_n1.json 2>&1; tail -1 gpurun_out/ev/train_n1.json
ls -la gpurun_out/ev
