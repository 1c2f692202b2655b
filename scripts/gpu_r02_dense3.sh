#!/bin/bash
O=gpurun_out/r02dense; mkdir -p $O
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:k_dense_tf32 -s 2 -c 1 -o $O/dense_fwd python scripts/learner_step_probe.py --backend tcgen05 --once --eager > $O/ncu_full.log 2>&1
ncu -i $O/dense_fwd.ncu-rep --page details > $O/k_dense_tf32_fwd_details.txt 2>/dev/null
grep -E "Duration|Throughput|Registers|Theoretical Occ|Achieved Occ|Waves|L2 Hit|Executed Ipc|Block Limit|Grid Size|Dynamic Shared" $O/k_dense_tf32_fwd_details.txt | head -40
