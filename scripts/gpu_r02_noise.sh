#!/bin/bash
# Decision-tree noise form: GPU parity suite, A/B timing against the all-vertices build, bench line, step ncu metrics
O=gpurun_out/r02noise; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
for rep in 1 2; do
  timeout 300 python scripts/step_timing.py --sizes 4096,8192,32768,65536 --variants fusedauto >> $O/step_timing_tree.jsonl 2>> $O/step_timing.err
  BLE_B200_LIB=$PWD/balloon_learning_environment_b200/variants/libble_allv.so timeout 300 python scripts/step_timing.py --sizes 4096,8192,32768,65536 --variants fusedauto >> $O/step_timing_allv.jsonl 2>> $O/step_timing.err
done
cat $O/step_timing_tree.jsonl $O/step_timing_allv.jsonl
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 1500 $O/bench_n1.json
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_step -c 12 --log-file $O/ncu_metrics_step.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu3.log 2>&1
tail -4 $O/ncu_metrics_step.csv
