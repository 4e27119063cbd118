#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02k_pytest.log
tail -3 gpurun_out/r02k_pytest.log
{
NRT_LANES=1 timeout 300 python tools/frame_breakdown.py config4
NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4 config3
NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4
} > gpurun_out/r02k_ab.log 2>&1
grep -v "fb sha" gpurun_out/r02k_ab.log
