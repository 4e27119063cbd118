#!/bin/bash
# session 3, call c: A/B of the by-value scene header in FusedBounce (libs under tools/ab/)
mkdir -p gpurun_out
{
for v in base byval hot; do
for rep in 1 2; do
echo "=== $v ($rep)"
NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
done
done
echo "=== hot, NRT_HOT_HEADER=0"
NRT_HOT_HEADER=0 NRT_LIB=/root/repo/tools/ab/libnrt_hot.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
echo "=== hot part 0,8"
NRT_PART=0,8 NRT_LIB=/root/repo/tools/ab/libnrt_hot.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
} > gpurun_out/s3c.log 2>&1
cut -c1-330 gpurun_out/s3c.log
