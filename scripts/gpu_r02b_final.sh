#!/bin/bash
# round 2, second session: evidence on the final library (decision-tree noise, 11-corner evaluation, tcgen05 dense layers,
# in-place host buffers): full GPU suite, smoke, bench lines, launch list, ncu metrics + full captures
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json
timeout 600 python bench.py --observation perciatelli --steps 10 --min-timed-ms 50 --no-cpu-baseline > $O/bench_n1_observation.json 2> $O/bench_n1_observation.err; tail -c 300 $O/bench_n1_observation.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; tail -c 400 $O/bench_reference_arm.json
timeout 300 python scripts/step_timing.py --sizes 4096,8192,16384,32768,65536 --variants fusedauto > $O/step_timing.jsonl 2>> $O/step_timing.err
timeout 900 ncu --csv --metrics gpu__time_duration.sum --clock-control none -c 400 --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_step -c 12 --log-file $O/ncu_metrics_step.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu3.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step_warp -s 6 -c 1 -o $O/step_warp python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > $O/bench_under_ncu4.log 2>&1
ncu -i $O/step_warp.ncu-rep --page details > $O/k_step_warp_65536_details.txt 2>/dev/null
ncu -i $O/step_warp.ncu-rep --page source --csv > $O/step_warp_source.csv 2>/dev/null; python scripts/ncu_source_report.py $O/step_warp_source.csv > $O/k_step_warp_65536_source_report.txt 2>/dev/null; rm -f $O/step_warp_source.csv
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy" $O/k_step_warp_65536_details.txt | head -6
# dense layers
for be in tcgen05 cublas; do timeout 300 python scripts/learner_step_probe.py --backend $be | tee -a $O/learner_step_probe.jsonl; done
timeout 300 python scripts/learner_step_probe.py --backend tcgen05 --eager | tee -a $O/learner_step_probe.jsonl
for be in tcgen05 cublas; do
  timeout 600 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --dense-backend $be 2> $O/train_$be.err | tail -1 > $O/train_qrdqn_n1_$be.json; cut -c1-120 $O/train_qrdqn_n1_$be.json; grep -o '"phase_ms_per_iteration.*' $O/train_qrdqn_n1_$be.json
done
for be in tcgen05 cublas; do
timeout 600 ncu --csv --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --log-file $O/launches_learner_step_$be.csv python scripts/learner_step_probe.py --backend $be --once --eager > $O/ncu_$be.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_dense_tf32 -s 10 -c 1 -o $O/dense_fwd python scripts/learner_step_probe.py --backend tcgen05 --once --eager > $O/ncu_full.log 2>&1
ncu -i $O/dense_fwd.ncu-rep --page details > $O/k_dense_tf32_details.txt 2>/dev/null
grep -E "k_dense|Duration|SM Active|Executed Ipc A" $O/k_dense_tf32_details.txt | head -6
rm -f $O/*.ncu-rep
