#!/bin/bash
mkdir -p gpurun_out
NRT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gate_write|k_for_each_counted|k_gate_flags|k_produce_gate" --launch-count 10 -o gpurun_out/r02_chain -f python tools/frame_breakdown.py config4 > gpurun_out/r02_chain_ncu.log 2>&1
tail -2 gpurun_out/r02_chain_ncu.log
