#!/bin/bash
# End-of-session evidence on one B200 (small outputs only: gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out/ev2
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/ev2/pytest_gpu.log 2>&1; tail -2 gpurun_out/ev2/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/ev2/bench_n1.json 2> gpurun_out/ev2/bench_n1.err; tail -c 200 gpurun_out/ev2/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/ev2/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --observation-probe 3 > gpurun_out/ev2/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gp_column4|k_gp_update|k_feat_range" -s 16 -c 4 -o gpurun_out/ev2/gp \
  python scripts/feature_timing.py --num-envs 16384 --fields 1024 > /dev/null 2>&1
ncu -i gpurun_out/ev2/gp.ncu-rep --page details > gpurun_out/ev2/gp_details.txt 2>/dev/null
rm -f gpurun_out/ev2/gp.ncu-rep
timeout 300 python scripts/feature_timing.py > gpurun_out/ev2/feature_timing.json 2>&1; tail -1 gpurun_out/ev2/feature_timing.json
timeout 300 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --warmup 130 > gpurun_out/ev2/train_n1.json 2>&1; tail -c 400 gpurun_out/ev2/train_n1.json
ls -la gpurun_out/ev2
