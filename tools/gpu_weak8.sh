#!/bin/bash
# N-GPU weak-scaling run of the default bench under torchrun (as the driver launches it)
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_weak_n$N.out 2> gpurun_out/r2_bench_weak_n$N.err; echo "bench weak N=$N rc=$?"
grep '^{' gpurun_out/r2_bench_weak_n$N.out | tail -1 > gpurun_out/r2_bench_weak_n$N.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_weak_n$N.json'))
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print(name, 'N', d['n_gpus'], r['scaling'], round(r['value']), 'e2e', round(r['e2e']['value']), 'ms', round(r['ms_per_step'],2), 'parity', r['parity'] and (r['parity']['checked'], r['parity']['bit_exact'], r['parity'].get('ranks_checked')))
show('a', d)
for k,v in (d.get('configs') or {}).items(): show(k, v)
PY
