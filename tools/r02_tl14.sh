#!/bin/bash
mkdir -p gpurun_out
{
for k in 0 7; do
echo "=== part $k,8"
NRT_TRACE_LANES=1 NRT_PART=$k,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha" | cut -c1-200
done
} > gpurun_out/r02zh.log 2>&1
cat gpurun_out/r02zh.log
