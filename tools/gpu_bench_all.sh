#!/bin/bash
# bench every BASELINE.json config on one GPU + launch lists; outputs under gpurun_out/
mkdir -p gpurun_out
for w in a f d prod16; do
  timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -2 gpurun_out/bench_$w.err
done
for w in f d; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ncu_launches_$w.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload $w --n 65536 > gpurun_out/ncu_launches_$w.out 2>&1; echo "ncu list $w rc=$?"
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_f_finalexp -s 1 -c 1 -o gpurun_out/prof_f_finalexp \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload f --n 37888 > gpurun_out/ncu_full_f.out 2>&1; echo "ncu full f rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_d_miller -s 1 -c 1 -o gpurun_out/prof_d_miller \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload d --n 37888 > gpurun_out/ncu_full_d.out 2>&1; echo "ncu full d rc=$?"
for w in a f d prod16; do python -c "
import json,sys
d=json.load(open('gpurun_out/bench_$w.json'))
print('$w', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['parity'], d['stage_ms'], 'frac', round(d['roofline']['frac'],3))
"; done
