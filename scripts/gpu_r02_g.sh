#!/bin/bash
# Round 2, call G: third-generation WindGP posterior (full refit per call: fp64 blocked Cholesky + 3xTF32 column sweep).
mkdir -p gpurun_out/r02g
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "features or incremental or cuda_balloon_arena or eval" > gpurun_out/r02g/pytest_feat.log 2>&1; grep -v "^$" gpurun_out/r02g/pytest_feat.log | tail -25
timeout 300 python scripts/feature_timing.py --num-envs 65536 2>&1 | tee gpurun_out/r02g/feature_timing.json | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_gp|k_feat" -s 100 -c 12 --csv --log-file gpurun_out/r02g/feat_launches.csv python scripts/feature_timing.py --num-envs 65536 > /dev/null 2>&1
grep -E "k_gp|k_feat" gpurun_out/r02g/feat_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
