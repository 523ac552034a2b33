#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_type_a.py tests/test_gpu_type_fd.py tests/test_gpu_shim.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --configs f,prod16 > gpurun_out/r2_bench_a9.json 2> gpurun_out/r2_bench_a9.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_a9.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_a9.json') if l.startswith('{')][-1])
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print(name, round(r['value']), 'e2e', round(r['e2e']['value']), 'same', r['e2e']['matches_device_resident_output'],
          'frac', round(r['roofline']['frac'],3), 'parity', r['parity'] and (r['parity']['checked'], r['parity']['bit_exact']), 'stage', {k: round(v,2) for k,v in r['stage_ms'].items()})
show('a', d)
for k,v in d['configs'].items(): show(k, v)
PY
