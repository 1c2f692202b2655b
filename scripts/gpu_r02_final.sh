#!/bin/bash
# round-2 evidence on the final library: full GPU suite, smoke, the bench lines, launch list, ncu facts
O=gpurun_out/r02final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json
timeout 600 python bench.py --observation perciatelli --steps 10 --min-timed-ms 50 --no-cpu-baseline > $O/bench_n1_observation.json 2> $O/bench_n1_observation.err; tail -c 300 $O/bench_n1_observation.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 400 $O/bench_reference_arm.json
timeout 900 ncu --csv --metrics gpu__time_duration.sum --clock-control none -c 400 --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_wind_gather -c 25 --log-file $O/ncu_metrics_gather.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu2.log 2>&1
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_step -c 12 --log-file $O/ncu_metrics_step.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_wind_gather -s 4 -c 1 -o $O/gather python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu4.log 2>&1
ncu -i $O/gather.ncu-rep --page details > $O/k_wind_gather_details.txt 2>/dev/null; rm -f $O/gather.ncu-rep
grep -E "Duration|DRAM Throughput|Memory Throughput" $O/k_wind_gather_details.txt | head -5
