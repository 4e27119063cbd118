#!/bin/bash
mkdir -p gpurun_out
{
for k in 0 1 2 3 4 5 6 7; do
echo "=== part $k,8"
NRT_TRACE_LANES=1 NRT_PART=$k,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | grep -v "^\[lanes\]" | head -3
NRT_TRACE_LANES=1 NRT_PART=$k,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep "^\[lanes\]" | tail -1
done
} > gpurun_out/r02zg.log 2>&1
cut -c1-250 gpurun_out/r02zg.log
