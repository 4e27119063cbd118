#!/bin/bash
mkdir -p gpurun_out
NRT_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err
tail -c 600 gpurun_out/r02g_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/r02g_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','gpu_launches','frames_per_s')}, d['e2e'], d.get('parity'))
        print([(k['kernel'][:20], round(k['ms'],3)) for k in d['kernels']])
PY
