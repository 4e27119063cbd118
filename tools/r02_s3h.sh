#!/bin/bash
# session 3, call h: per-launch prefilter trace (rays, runs, work items, time) of a whole frame and of a 1/8 frame, one lane
mkdir -p gpurun_out
{
echo "=== FULL one lane"
NRT_LANES=1 NRT_TRACE_PREFILTER=1 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -E "prefilter\]|config4:" | tail -12
echo "=== part 0,8 one lane"
NRT_LANES=1 NRT_PART=0,8 NRT_TRACE_PREFILTER=1 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -E "prefilter\]|config4:" | tail -12
} > gpurun_out/s3h.log 2>&1
cut -c1-260 gpurun_out/s3h.log
