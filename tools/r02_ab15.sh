#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k config5_class 2>&1 | tail -40
NRT_LANES=4 NRT_PART=0,8 timeout 600 python tools/frame_breakdown.py config5 2>&1 | grep -v "fb sha"
