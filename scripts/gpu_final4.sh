#!/bin/bash
mkdir -p gpurun_out/final4
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_step or errors_and_host" > gpurun_out/final4/pytest_host.log 2>&1; tail -15 gpurun_out/final4/pytest_host.log
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/final4/bench_n1.json 2> gpurun_out/final4/bench_n1.err; tail -c 100 gpurun_out/final4/bench_n1.json
python - <<'PY'
import json
for l in open('gpurun_out/final4/bench_n1.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'] / 1e6, 'e2e', d['e2e']['value'] / 1e6, 'launches', d['gpu_launches'])
PY
