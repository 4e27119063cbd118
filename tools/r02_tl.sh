#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tl_*.txt
{
NRT_TIMELINE=gpurun_out/tl_part8.txt NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
NRT_TIMELINE=gpurun_out/tl_full.txt NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4
NRT_TRACE_PREFILTER=1 NRT_LANES=1 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
} > gpurun_out/r02u.log 2>&1
grep -v "fb sha" gpurun_out/r02u.log | tail -60
