#!/bin/bash
# The round's `ncu --set full --import-source on` captures (one kernel launch each), reduced to the text files under profiles/:
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r02_profiles.sh'   ->  gpurun_out/r02prof/*_details.txt, *_source.csv
# `python scripts/ncu_source_report.py <source.csv>` buckets the stall samples along the SASS.
O=gpurun_out/r02prof; mkdir -p $O
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$regex -s $skip -c 1 -o $O/$name "$@" > $O/$name.log 2>&1
  ncu -i $O/$name.ncu-rep --page details > $O/${name}_details.txt 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv > $O/${name}_source.csv 2>/dev/null
  rm -f $O/$name.ncu-rep
  grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Achieved Occupancy" $O/${name}_details.txt | head -4
}
cap k_step_warp_65536 k_step_warp 20 python scripts/step_timing.py --sizes 65536 --variants fused0 --steps 40
cap k_step_roles8_8192 k_step_roles 20 python scripts/step_timing.py --sizes 8192 --variants fused8 --steps 40
cap k_gp_posterior_16384 k_gp_posterior 2 python scripts/feature_timing.py --num-envs 16384 --fields 1024
cap k_flow_to_windows_2048 k_flow_to_windows 3 python scripts/gen_timing.py --fields 8192 --reps 1
