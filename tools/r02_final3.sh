#!/bin/bash
# end-of-round-2 measurement set on one B200 (the .ncu-rep stays on the box: gpurun_out/ is capped at 64 MiB)
mkdir -p gpurun_out
NRT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_for_each_stats|k_mesh_prefilter|k_prefilter_bounds|k_gate_write|k_produce_gate|k_gate_flags|k_finalize|k_path_warp|k_for_each_counted|k_sel_" --launch-count 30 -o /tmp/r02b_frame -f python tools/frame_breakdown.py config4 > gpurun_out/r02b_ncu_full.log 2>&1; echo "ncu full rc $?"
python tools/make_kernel_traffic.py /tmp/r02b_frame.ncu-rep config4 132710400 profiles/kernel_traffic.json > gpurun_out/r02b_kernel_traffic.log 2>&1; echo "traffic rc $?"
cp profiles/kernel_traffic.json gpurun_out/kernel_traffic.json
python tools/ncu_metrics.py /tmp/r02b_frame.ncu-rep 12 > gpurun_out/r02b_ncu_frame_config4.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_config4_n1.json 2> gpurun_out/r02b_bench_config4_n1.err; echo "bench rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02b_launches_bench_config4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu_launches.log 2>&1; echo "ncu launches rc $?"
du -sh gpurun_out
