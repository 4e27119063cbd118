#!/bin/bash
mkdir -p gpurun_out
{
for v in 3 4; do NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_occ$v.so NRT_LANES=1 timeout 300 python tools/frame_breakdown.py config4; done
NRT_LANES=1 timeout 300 python tools/frame_breakdown.py config4
for v in 3 4; do NRT_LIB=$PWD/nim_raytracer_b200/csrc/libnrt_occ$v.so NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4; done
} > gpurun_out/r02j_ab.log 2>&1
grep -v "fb sha" gpurun_out/r02j_ab.log
