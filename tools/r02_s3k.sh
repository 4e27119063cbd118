#!/bin/bash
# session 3, call k: A/B of the per-lane-mask survivor emission in k_mesh_prefilter
mkdir -p gpurun_out
{
for rep in 1 2; do
for v in head emit; do
echo "=== $v ($rep)"
NRT_LANES=1 NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
done
done
for v in head emit; do
echo "=== $v 4 lanes; part 0,8; config 3/2"
NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
NRT_PART=0,8 NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config3 config2 2>&1 | grep -v "active/bounce"
done
} > gpurun_out/s3k.log 2>&1
cut -c1-330 gpurun_out/s3k.log
NRT_LIB=/root/repo/tools/ab/libnrt_emit.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
