#!/bin/bash
# 11 lower-frame corners (main) vs 16 masked corners: parity suite on main, A/B timing, ncu instruction count.
# variants/libble_c16.so = the library of commit e32193b (tree form, 16 masked corners); the all-vertices A/B library is
# `scripts/build_variant.sh allv -DBLE_NOISE_ALL_VERTICES` (substitute it for libble_c16.so to reproduce r02b_step_timing_noise_all_vertices.jsonl)
O=gpurun_out/r02noise2; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
for rep in 1 2; do
  timeout 300 python scripts/step_timing.py --sizes 4096,8192,32768,65536 --variants fusedauto >> $O/step_timing_c11.jsonl 2>> $O/step_timing.err
  BLE_B200_LIB=$PWD/balloon_learning_environment_b200/variants/libble_c16.so timeout 300 python scripts/step_timing.py --sizes 4096,8192,32768,65536 --variants fusedauto >> $O/step_timing_c16.jsonl 2>> $O/step_timing.err
done
cat $O/step_timing_c11.jsonl $O/step_timing_c16.jsonl | cut -c1-140
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_step -c 12 --log-file $O/ncu_metrics_step.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu3.log 2>&1
grep smsp__inst $O/ncu_metrics_step.csv | tail -2
