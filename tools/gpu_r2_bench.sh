#!/bin/bash
# round 2: the default bench run (all BASELINE configs) + the GPU test suite on the current build
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_default.json') if l.startswith('{')][-1])
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print(name, round(r['value']), 'e2e', round(r['e2e']['value']), 'same', r['e2e']['matches_device_resident_output'],
          'frac', round(r['roofline']['frac'],3), 'parity', r['parity'], 'stage', {k: round(v,2) for k,v in r['stage_ms'].items()},
          'cpu', r['cpu_baseline'] and round(r['cpu_baseline']['value']))
show('a', d)
for k,v in d['configs'].items(): show(k, v)
print('fp_mul', json.dumps(d['fp_mul'])[:600])
print('bench.c', d['reference_benchmark_c'])
PY
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
fi
