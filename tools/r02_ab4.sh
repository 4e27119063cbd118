#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_for_each_stats --launch-count 1 -o gpurun_out/r02d_fused -f python tools/frame_breakdown.py config4 > gpurun_out/r02d_ncu.log 2>&1
tail -3 gpurun_out/r02d_ncu.log
NRT_PART=0,8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02d_launches_part8.csv python tools/frame_breakdown.py config4 > gpurun_out/r02d_ncu2.log 2>&1
tail -2 gpurun_out/r02d_ncu2.log
ls -la gpurun_out/
