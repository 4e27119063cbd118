#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02i_pytest.log
tail -3 gpurun_out/r02i_pytest.log
{
for t in 1 4; do NRT_TILE=$t NRT_LANES=1 timeout 300 python tools/frame_breakdown.py config4; done
for t in 1 4; do NRT_TILE=$t NRT_LANES=4 NRT_PART=0,8 timeout 300 python tools/frame_breakdown.py config4; done
NRT_LANES=4 timeout 300 python tools/frame_breakdown.py config4 config3 config2
} > gpurun_out/r02i_ab.log 2>&1
cat gpurun_out/r02i_ab.log
