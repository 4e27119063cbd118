#!/bin/bash
mkdir -p gpurun_out
{
for sp in 0 2 8 32 128; do
echo "=== SPLIT $sp"
NRT_PREFILTER_SPLIT=$sp NRT_TRACE_PREFILTER=1 NRT_LANES=1 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | tail -14
done
for ln in 1 2 3 4; do
echo "=== LANES $ln"
NRT_LANES=$ln NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "fb sha"
done
} > gpurun_out/r02v.log 2>&1
grep -v "fb sha" gpurun_out/r02v.log | cut -c1-330
