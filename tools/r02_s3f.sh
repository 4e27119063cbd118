#!/bin/bash
# session 3, call f: A/B of the 16-flags-per-thread ordered select against cub::DeviceSelect
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in sel sel2; do
echo "=== $v ($rep)"
NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
done
done
echo "=== sel part 0,8 + others"
NRT_PART=0,8 NRT_LIB=/root/repo/tools/ab/libnrt_sel2.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
NRT_LIB=/root/repo/tools/ab/libnrt_sel2.so timeout 300 python tools/frame_breakdown.py config3 config2 config1 2>&1 | grep -v "active/bounce"
} > gpurun_out/s3g.log 2>&1
cut -c1-330 gpurun_out/s3g.log
