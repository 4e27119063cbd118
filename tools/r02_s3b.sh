#!/bin/bash
# session 3, call b: source-level ncu capture of FusedBounce's bounce-0 launch (config 4, one lane)
mkdir -p gpurun_out
NRT_LANES=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_for_each_stats" --launch-count 1 -o gpurun_out/s3b_fb -f python tools/frame_breakdown.py config4 > gpurun_out/s3b_ncu.log 2>&1
tail -2 gpurun_out/s3b_ncu.log
ls -la gpurun_out/s3b_fb.ncu-rep
