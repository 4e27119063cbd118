#!/bin/bash
mkdir -p gpurun_out
NRT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mesh_prefilter --launch-count 3 -o gpurun_out/r02_prefilter -f python tools/frame_breakdown.py config4 > gpurun_out/r02_prefilter_ncu.log 2>&1
tail -2 gpurun_out/r02_prefilter_ncu.log
ls -la gpurun_out/r02_prefilter.ncu-rep
