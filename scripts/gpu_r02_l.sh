#!/bin/bash
# generation path: fused flow->windows writer, TF32 decoder, KS tests, real checkpoint; then bench + ncu facts
mkdir -p gpurun_out/r02l
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "decoder or fused_field or latents or reset" > gpurun_out/r02l/pytest_gen.log 2>&1; grep -E "passed|failed|error|KS|decoder|real checkpoint" gpurun_out/r02l/pytest_gen.log | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02l/bench_n1.json 2> gpurun_out/r02l/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','gpu_launches_per_step')}); print(d['e2e']['value']); print(d['reset_path'])
PY
timeout 900 ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'k_wind_gather|k_step_warp|k_step_roles' -c 60 --log-file gpurun_out/r02l/ncu_facts.csv python bench.py --steps 2 --warmup 3 --min-timed-ms 0.5 --no-cpu-baseline > gpurun_out/r02l/bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02l/ncu_facts.csv | cut -c1-300
timeout 600 ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_flow_to_windows|sgemm|gemm|k_sample_latents|k_reset|k_make_perms|Kernel|cutlass' -c 80 --log-file gpurun_out/r02l/ncu_gen.csv python - > gpurun_out/r02l/gen_under_ncu.log 2>&1 <<'PY'
import torch
from balloon_learning_environment_b200 import batched_env, models
n = 16384
a = batched_env.BatchedBalloonArena(n, precision='fp32', field_layout='x64')
a.set_decoder(models.load_decoder(''))
a.alloc_wind_fields(n)
seeds = torch.arange(n, dtype=torch.int64)
a.sample_wind_fields(seeds); torch.cuda.synchronize()
a.sample_wind_fields(seeds + 7); torch.cuda.synchronize()
PY
tail -2 gpurun_out/r02l/ncu_gen.csv | cut -c1-300
