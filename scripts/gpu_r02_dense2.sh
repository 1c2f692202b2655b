#!/bin/bash
O=gpurun_out/r02dense; mkdir -p $O
for be in tcgen05 cublas; do timeout 300 python scripts/learner_step_probe.py --backend $be | tee -a $O/learner_step_probe.jsonl; done
timeout 300 python scripts/learner_step_probe.py --backend tcgen05 --eager | tee -a $O/learner_step_probe.jsonl
for be in tcgen05 cublas; do
timeout 600 ncu --csv --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --log-file $O/launches_learner_step_$be.csv python scripts/learner_step_probe.py --backend $be --once --eager > $O/ncu_$be.log 2>&1
python - <<PY
import csv, collections
rows = [l for l in open('$O/launches_learner_step_$be.csv') if l.startswith('"')]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(rows):
  k = r['Kernel Name'][:70]; tot[k][0] += 1; tot[k][1] += float(r['Metric Value'].replace(',', '')) / 1e3
print('$be', sum(v[1] for v in tot.values()), 'us total')
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]: print(f'  {v[1]:9.1f} us  x{v[0]:3d}  {k}')
PY
done
