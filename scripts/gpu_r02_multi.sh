#!/bin/bash
# N>1 evidence on one 8-GPU box: strong scaling of BASELINE configs[2] (65,536 balloons split over N GPUs; the weak number is a
# side key of the same line), and the product path that broadcasts the decoder weights over NCCL at construction (run_eval).
mkdir -p gpurun_out/r02multi
for N in 2 4 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 500 $TR --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 100 > gpurun_out/r02multi/bench_n${N}.json 2> gpurun_out/r02multi/bench_n${N}.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/r02multi/bench_n${N}.json').read().strip().splitlines()[-1])
  print($N, d['value'], d['ms_per_step'], d['scaling'], d.get('rollout',{}).get('value'), d.get('weak_scaling',{}).get('value'), d.get('weak_scaling',{}).get('rollout_value'))
except Exception as e:
  print('bench N=$N failed', e); print(open('gpurun_out/r02multi/bench_n${N}.err').read()[-1500:])
PY
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 500 $TR --master-port 29541 scripts/run_eval.py --agent station_seeker --suite small_eval --max-episode-length 120 --no-flight-path > gpurun_out/r02multi/run_eval_n2.log 2>&1
grep -E "env_steps_per_s|NVLS|Broadcast|broadcast|Error|error" gpurun_out/r02multi/run_eval_n2.log | cut -c1-400 | tail -8
