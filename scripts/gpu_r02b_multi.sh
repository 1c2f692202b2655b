#!/bin/bash
# 8-GPU evidence on the final library: strong scaling of BASELINE configs[2] (weak number in the same line), N = 1 on the
# same box for the ratio, configs[3] on 8 GPUs, configs[4] (QR-DQN training, tcgen05 dense path) on 8 GPUs
O=gpurun_out/r02bmulti; mkdir -p $O
timeout 400 python bench.py --steps 20 --warmup 5 --min-timed-ms 100 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
python -c "
import json; d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1]); print(1, d['value'], d['ms_per_step'], d['rollout']['value'])"
for N in 8; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 500 $TR --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 --min-timed-ms 100 > $O/bench_n${N}.json 2> $O/bench_n${N}.err
  python - <<PY
import json
try:
  d=json.loads(open('$O/bench_n${N}.json').read().strip().splitlines()[-1])
  print($N, d['value'], d['ms_per_step'], d['scaling'], d.get('rollout',{}).get('value'), d.get('weak_scaling',{}).get('value'), d.get('weak_scaling',{}).get('rollout_value'))
except Exception as e:
  print('bench N=$N failed', e); print(open('$O/bench_n${N}.err').read()[-1500:])
PY
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29551 bench.py --gpus 8 --num-envs 32768 --observation perciatelli --steps 10 --warmup 3 --min-timed-ms 50 --scaling weak > $O/bench_obs_n8.json 2> $O/bench_obs_n8.err
python -c "
import json; d=json.loads(open('$O/bench_obs_n8.json').read().strip().splitlines()[-1]); print('obs8', d['value'], d['ms_per_step'], d['config']['num_envs'], d['e2e']['value'])"
timeout 500 $TR --master-port 29552 scripts/train_qrdqn.py --num-envs 32768 --iterations 30 2> $O/train_n8.err | tail -1 > $O/train_qrdqn_n8_tcgen05.json; cut -c1-700 $O/train_qrdqn_n8_tcgen05.json
