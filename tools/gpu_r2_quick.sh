#!/bin/bash
# short measurement pass: GPU tests, smoke, the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | grep real
tail -2 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_default.json') if l.startswith('{')][-1])
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print(name, round(r['value']), 'e2e', round(r['e2e']['value']), 'same', r['e2e']['matches_device_resident_output'],
          'frac', round(r['roofline']['frac'],3), 'parity', r['parity'] and (r['parity']['checked'], r['parity']['bit_exact']), 'stage', {k: round(v,2) for k,v in r['stage_ms'].items()})
show('a', d)
for k,v in (d.get('configs') or {}).items(): show(k, v)
PY
