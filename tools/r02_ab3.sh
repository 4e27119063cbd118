#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
{
NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4 config3 config2
NRT_PATH=1 NRT_TAIL_BELOW=0 timeout 300 python tools/frame_breakdown.py config4
NRT_PATH=1 NRT_TAIL_BELOW=1000000 timeout 300 python tools/frame_breakdown.py config4
for part in 0,8 3,8; do NRT_PART=$part NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4; NRT_PART=$part NRT_TAIL_BELOW=0 NRT_PATH=1 timeout 300 python tools/frame_breakdown.py config4; done
} > gpurun_out/r02c_ab.log 2>&1
cat gpurun_out/r02c_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:FusedBounce --launch-count 1 -o gpurun_out/r02c_fused -f python tools/frame_breakdown.py config4 > gpurun_out/r02c_ncu.log 2>&1
tail -3 gpurun_out/r02c_ncu.log
ls -la gpurun_out/
