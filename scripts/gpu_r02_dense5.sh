#!/bin/bash
# tcgen05 dense path: tests, SGD-step probe (both backends), training loop A/B, launch list of one step, ncu --set full of one forward product
O=gpurun_out/r02dense_final; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_learner.py -m gpu -q 2>&1 | tail -4 | tee $O/pytest_dense.log
for be in tcgen05 cublas; do timeout 300 python scripts/learner_step_probe.py --backend $be | tee -a $O/learner_step_probe.jsonl; done
timeout 300 python scripts/learner_step_probe.py --backend tcgen05 --eager | tee -a $O/learner_step_probe.jsonl
for be in tcgen05 cublas; do
  timeout 600 python scripts/train_qrdqn.py --num-envs 4096 --iterations 30 --dense-backend $be 2> $O/train_$be.err | tail -1 | tee $O/train_qrdqn_n1_$be.json
done
be=tcgen05
timeout 600 ncu --csv --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --log-file $O/launches_learner_step_$be.csv python scripts/learner_step_probe.py --backend $be --once --eager > $O/ncu_$be.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_dense_tf32 -s 10 -c 1 -o $O/dense_fwd python scripts/learner_step_probe.py --backend tcgen05 --once --eager > $O/ncu_full.log 2>&1
ncu -i $O/dense_fwd.ncu-rep --page details > $O/k_dense_tf32_details.txt 2>/dev/null
grep -E "k_dense|Duration|Throughput|SM Active|Waves|Executed Ipc" $O/k_dense_tf32_details.txt | head -14
