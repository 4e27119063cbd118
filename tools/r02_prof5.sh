#!/bin/bash
mkdir -p gpurun_out
NRT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_for_each_stats" --launch-count 1 -o gpurun_out/r02_fb2 -f python tools/frame_breakdown.py config4 > gpurun_out/r02_fb2_ncu.log 2>&1
tail -2 gpurun_out/r02_fb2_ncu.log
