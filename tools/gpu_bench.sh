#!/bin/bash
# round-1 measurement pass on the GPU box: tests, bench (both arms), ubench, ncu launch list + full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_b200.json 2> gpurun_out/bench_b200.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
./build/ubench > gpurun_out/ubench.jsonl 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ncu_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --n 262144 > gpurun_out/ncu_launches.out 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_a_miller -s 1 -c 1 -o gpurun_out/prof_a_miller \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --n 151552 > gpurun_out/ncu_full.out 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_b200.json; tail -3 gpurun_out/bench_b200.err; cat gpurun_out/bench_reference.json; cat gpurun_out/ubench.jsonl; tail -5 gpurun_out/ncu_launches.csv; tail -3 gpurun_out/ncu_full.out
