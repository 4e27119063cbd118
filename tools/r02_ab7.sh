#!/bin/bash
mkdir -p gpurun_out
NRT_LANES=1 NRT_PART=0,8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mesh_prefilter --launch-count 3 -o gpurun_out/r02h_prefilter_part8 -f python tools/frame_breakdown.py config4 > gpurun_out/r02h_ncu.log 2>&1
tail -2 gpurun_out/r02h_ncu.log
NRT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mesh_prefilter --launch-count 1 -o gpurun_out/r02h_prefilter_full -f python tools/frame_breakdown.py config4 > gpurun_out/r02h_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
