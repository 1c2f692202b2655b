#!/bin/bash
mkdir -p gpurun_out/r02n
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decoder or fused_field or latents or eval or reset" > gpurun_out/r02n/pytest_gen.log 2>&1; tail -3 gpurun_out/r02n/pytest_gen.log
for l in x64 x128; do timeout 300 python scripts/gen_timing.py --fields 32768 --layout $l | tee -a gpurun_out/r02n/gen_timing.jsonl; done
timeout 300 python scripts/gen_timing.py --fields 65536 | tee -a gpurun_out/r02n/gen_timing.jsonl
timeout 300 python scripts/gen_timing.py --fields 4096 --reps 10 | tee -a gpurun_out/r02n/gen_timing.jsonl
timeout 600 ncu --csv --metrics gpu__time_duration.sum --clock-control none -c 60 --log-file gpurun_out/r02n/ncu_gen.csv python scripts/gen_timing.py --fields 8192 --reps 1 > gpurun_out/r02n/gen_under_ncu.log 2>&1
