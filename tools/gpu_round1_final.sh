#!/bin/bash
# refresh every measurement with the current build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shim.py -q --timeout 600 2>&1 | tail -3
timeout 600 python tools/gpu_group_bench.py 262144 > gpurun_out/group_bench.jsonl 2> gpurun_out/group_bench.err; echo "group bench rc=$?"; cat gpurun_out/group_bench.jsonl; tail -2 gpurun_out/group_bench.err
for w in a f d prod16; do
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -2 gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2>/dev/null
for w in a f d prod16; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][-1])
print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['parity'], d['stage_ms'], 'frac', round(d['roofline']['frac'],3))
"; done
