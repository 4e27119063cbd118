#!/bin/bash
# session 3, call j: source-level capture of the prefilter launches (ORIGIN, DIR, DIR) of a config-4 frame, one lane
mkdir -p gpurun_out
NRT_LANES=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_mesh_prefilter" --launch-count 4 -o gpurun_out/s3j_pre -f python tools/frame_breakdown.py config4 > gpurun_out/s3j_ncu.log 2>&1
tail -1 gpurun_out/s3j_ncu.log
