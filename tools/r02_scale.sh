#!/bin/bash
# round-2 scaling set on one 8 x B200 box: bench.py at N = 2, 4, 8 (torchrun, one process per GPU), the one-process
# 8-GPU mode, and BASELINE config 5 on eight
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_config4_n1.json 2>/dev/null; echo "n1 rc $?"
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02_scale_config4_n$n.json 2> gpurun_out/r02_scale_config4_n$n.err; echo "n$n rc $?"
done
timeout 600 python bench.py --gpus 8 --inproc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_config4_inproc8.json 2> gpurun_out/r02_bench_config4_inproc8.err; echo "inproc rc $?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --workload config5 --steps 5 --warmup 3 > gpurun_out/r02_bench_config5_n8.json 2> gpurun_out/r02_bench_config5_n8.err; echo "c5 rc $?"
python -m pytest tests/test_gpu_parity.py -m gpu -q -k two_gpus 2>&1 | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_scale_config4_n*.json'))+['gpurun_out/r02_bench_config4_inproc8.json','gpurun_out/r02_bench_config5_n8.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['n_gpus'], round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'ms', (d.get('parity') or {}).get('matches_oracle'))
PY
tail -3 gpurun_out/r02_bench_config4_inproc8.err gpurun_out/r02_bench_config5_n8.err
