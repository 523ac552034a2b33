#!/bin/bash
# final measurement pass of round 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3
show() { python - "$1" <<'PY'
import json, sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print(name, round(r['value']), 'e2e', round(r['e2e']['value']), 'same', r['e2e']['matches_device_resident_output'],
          'frac', round(r['roofline']['frac'],3), 'parity', r['parity'] and (r['parity']['checked'], r['parity']['bit_exact']), 'stage', {k: round(v,2) for k,v in r['stage_ms'].items()})
show('a', d)
for k,v in (d.get('configs') or {}).items(): show(k, v)
PY
}
( time timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err ) 2>&1 | grep real
tail -2 gpurun_out/r2_bench_default.err; show gpurun_out/r2_bench_default.json
( time timeout 1500 python bench.py --steps 5 --warmup 3 --cpu-seconds 70 > gpurun_out/r2_bench_fullparity.json 2> gpurun_out/r2_bench_fullparity.err ) 2>&1 | grep real
tail -2 gpurun_out/r2_bench_fullparity.err; show gpurun_out/r2_bench_fullparity.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference.json
for w in pp g a1; do timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --configs none > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"; show gpurun_out/r2_bench_$w.json; done
timeout 600 python tools/gpu_pp_compare.py a f d g 131072 > gpurun_out/r2_pp_compare.jsonl 2> gpurun_out/r2_pp_compare.err; cut -c1-200 gpurun_out/r2_pp_compare.jsonl
PBC_B200_UNDER_PROFILER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_ncu_launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none --n 262144 > gpurun_out/ncu_launches_a.out 2>&1; echo "ncu list a rc=$?"
PBC_B200_UNDER_PROFILER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_ncu_launches_f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload f --configs none --n 262144 > gpurun_out/ncu_launches_f.out 2>&1; echo "ncu list f rc=$?"
KERNELS="${KERNELS:-k_a_miller9:a:227328 k_f_miller_s:f:151552 k_f_finalexp_s:f:151552 k_a_finalexp:a:227328}" bash tools/gpu_ncu.sh
