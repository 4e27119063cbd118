#!/bin/bash
# session 3, call e: A/B of the cx/cy tables and the field-wise light reads
mkdir -p gpurun_out
{
for rep in 1 2; do
echo "=== all ($rep)"
NRT_LIB=/root/repo/tools/ab/libnrt_all.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
echo "=== tab, tables off ($rep)"
NRT_GEN_TABLES=0 NRT_LIB=/root/repo/tools/ab/libnrt_tab.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
echo "=== tab ($rep)"
NRT_LIB=/root/repo/tools/ab/libnrt_tab.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
done
echo "=== tab part 0,8 + others"
NRT_PART=0,8 NRT_LIB=/root/repo/tools/ab/libnrt_tab.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
NRT_LIB=/root/repo/tools/ab/libnrt_tab.so timeout 300 python tools/frame_breakdown.py config3 config2 config1 2>&1 | grep -v "active/bounce"
} > gpurun_out/s3e.log 2>&1
cut -c1-330 gpurun_out/s3e.log
