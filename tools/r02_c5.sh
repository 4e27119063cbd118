#!/bin/bash
mkdir -p gpurun_out
{
NRT_LANES=1 timeout 600 python tools/frame_breakdown.py config5s
NRT_LANES=4 NRT_PART=0,8 timeout 600 python tools/frame_breakdown.py config5
} > gpurun_out/r02p_c5.log 2>&1
cat gpurun_out/r02p_c5.log
