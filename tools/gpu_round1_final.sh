#!/bin/bash
# refresh every measurement with the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 600 python tools/gpu_group_bench.py 262144 > gpurun_out/group_bench.jsonl 2> gpurun_out/group_bench.err; echo "group bench rc=$?"; cut -c1-200 gpurun_out/group_bench.jsonl; tail -2 gpurun_out/group_bench.err
for w in a f d prod16; do
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $w --cpu-seconds 6 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -2 gpurun_out/bench_$w.err
done
for w in a f d prod16; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_$w.json') if l.startswith('{')][-1])
print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['parity'], d['stage_ms'], 'frac', round(d['roofline']['frac'],3), d['roofline'].get('frac_executed'))
"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ncu_launches_a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --n 262144 > gpurun_out/ncu_launches_a.out 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_a_miller -s 1 -c 1 -o gpurun_out/prof_a_miller_w12 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --n 151552 > gpurun_out/ncu_full_a.out 2>&1; echo "ncu full a rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_f_miller -s 1 -c 1 -o gpurun_out/prof_f_miller_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload f --n 75776 > gpurun_out/ncu_full_f.out 2>&1; echo "ncu full f rc=$?"
