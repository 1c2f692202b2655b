#!/bin/bash
mkdir -p gpurun_out/r02aa
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02aa/pytest.log 2>&1; tail -3 gpurun_out/r02aa/pytest.log
python - <<'PY'
import torch, time
from balloon_learning_environment_b200 import batched_env
n=65536
a=batched_env.BatchedBalloonArena(n, precision='fp32', wind_model='simple_static', enable_noise=True)
seeds=torch.arange(n,dtype=torch.int64)
a.reset(seeds); torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(5): a.reset(seeds+r)
e1.record(); torch.cuda.synchronize()
print('reset_ms', e0.elapsed_time(e1)/5)
PY
