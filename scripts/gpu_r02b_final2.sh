#!/bin/bash
# final library of the round: full GPU suite, smoke, bench line (now with the qrdqn_sgd_step key), learner evidence refreshed
O=gpurun_out/r02b2; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 900 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 300 python scripts/step_timing.py --sizes 4096,8192,16384,32768,65536 --variants fusedauto > $O/step_timing.jsonl 2>> $O/step_timing.err
for be in tcgen05 cublas; do timeout 300 python scripts/learner_step_probe.py --backend $be | tee -a $O/learner_step_probe.jsonl; done
timeout 300 python scripts/learner_step_probe.py --backend tcgen05 --eager | tee -a $O/learner_step_probe.jsonl
for be in tcgen05 cublas; do
  timeout 600 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --dense-backend $be 2> $O/train_$be.err | tail -1 > $O/train_qrdqn_n1_$be.json; grep -o '"ms_per_iteration.\{0,25\}\|"phase_ms_per_iteration.*' $O/train_qrdqn_n1_$be.json
done
timeout 600 ncu --csv --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --log-file $O/launches_learner_step_tcgen05.csv python scripts/learner_step_probe.py --backend tcgen05 --once --eager > $O/ncu_tcgen05.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_dense_tf32 -s 10 -c 1 -o $O/dense_fwd python scripts/learner_step_probe.py --backend tcgen05 --once --eager > $O/ncu_full.log 2>&1
ncu -i $O/dense_fwd.ncu-rep --page details > $O/k_dense_tf32_details.txt 2>/dev/null
rm -f $O/*.ncu-rep
