#!/bin/bash
mkdir -p gpurun_out/r02h
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_gp_posterior|k_feat_range|k_feat_ambient" -c 12 --csv --log-file gpurun_out/r02h/feat_launches.csv python scripts/feature_timing.py --num-envs 65536 > /dev/null 2>&1
grep -E "k_gp|k_feat" gpurun_out/r02h/feat_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gp_posterior -s 2 -c 1 -o gpurun_out/r02h/gp python scripts/feature_timing.py --num-envs 16384 --fields 1024 > gpurun_out/r02h/ncu.log 2>&1
ncu -i gpurun_out/r02h/gp.ncu-rep --page details > gpurun_out/r02h/gp_details.txt 2>/dev/null
ncu -i gpurun_out/r02h/gp.ncu-rep --page source --csv > gpurun_out/r02h/gp_source.csv 2>/dev/null
rm -f gpurun_out/r02h/gp.ncu-rep
grep -E "Duration|Executed Ipc|Issue Slots Busy|Active Warps Per|Warp Cycles Per Issued|Executed Instructions|Registers Per|Achieved Active|Block Limit|Theoretical Occ" gpurun_out/r02h/gp_details.txt
