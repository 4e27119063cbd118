#!/bin/bash
# session 3, call d: GPU tests on the by-value build, A/B hot (FusedBounce only) vs all (every per-sample functor), source-level capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
{
for v in hot all; do
for rep in 1 2; do
echo "=== $v ($rep)"
NRT_LIB=/root/repo/tools/ab/libnrt_$v.so timeout 300 python tools/frame_breakdown.py config4 2>&1 | grep -v "active/bounce"
done
done
echo "=== all part 0,8"
NRT_PART=0,8 NRT_LIB=/root/repo/tools/ab/libnrt_all.so timeout 300 python tools/frame_breakdown.py config4 config3 config2 2>&1 | grep -v "active/bounce"
} > gpurun_out/s3d.log 2>&1
cut -c1-330 gpurun_out/s3d.log
NRT_LANES=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_for_each_stats" --launch-count 1 -o gpurun_out/s3d_fb -f python tools/frame_breakdown.py config4 > gpurun_out/s3d_ncu.log 2>&1
tail -1 gpurun_out/s3d_ncu.log
