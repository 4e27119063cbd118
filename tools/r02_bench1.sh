#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err; echo "bench rc $?"
tail -c 1500 gpurun_out/r02l_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/r02l_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','gpu_launches','frames_per_s','frame_hbm_bytes_model')})
        print('e2e', d['e2e']); print('parity', d['parity']); print('roofline', json.dumps(d['roofline'])[:1500]); print('intersection', {k:d['intersection_kernel'][k] for k in ('achieved','frac','achieved_alone','frac_alone','tests_per_step')})
        print('fused', d['fused_path']); print('cpu', d.get('cpu_baseline')); print('clocks', d['clocks'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
