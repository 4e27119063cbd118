#!/bin/bash
mkdir -p gpurun_out
{
for v in 4 5; do
NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_pre$v.so NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_pre$v.so NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4
done
NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
} > gpurun_out/r02s.log 2>&1
grep -v "fb sha\|active/" gpurun_out/r02s.log
