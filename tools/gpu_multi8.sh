#!/bin/bash
# N-GPU pass (gpurun --gpus N): weak and strong scaling of the default bench under torchrun (as the driver
# launches it), then the single-process fan-out (pbc_b200_set_devices) on the 8 x 2^20 weak batch
mkdir -p gpurun_out
N=${NGPU:-8}
run() {   # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 $2 > gpurun_out/r2_bench_$1_n$N.out 2> gpurun_out/r2_bench_$1_n$N.err; echo "bench $1 N=$N rc=$?"
  grep '^{' gpurun_out/r2_bench_$1_n$N.out | tail -1 > gpurun_out/r2_bench_$1_n$N.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_$1_n$N.json'))
def show(name, r):
    if 'error' in r: print(name, 'ERROR', r['error']); return
    print('$1', name, 'N', d['n_gpus'], r['scaling'], round(r['value']), 'e2e', round(r['e2e']['value']), 'ms', round(r['ms_per_step'],2), 'parity', r['parity'] and (r['parity']['checked'], r['parity']['bit_exact'], r['parity']['ranks_checked']))
show('a', d)
for k,v in (d.get('configs') or {}).items(): show(k, v)
PY
}
run weak ""
run strong "--scaling strong --configs prod16"
# one process driving all N devices through the C ABI
timeout 900 python - <<PY
import json, time, torch
import bench
from pbc_b200.pairing import Pairing
from pbc_b200.params import PARAMS
w = bench.WORKLOADS["a"]; n = $N << 20
P, Q = bench.make_inputs(w, n)
pr = Pairing(PARAMS["a"])
Pp, Qp = torch.from_numpy(P.copy()).pin_memory(), torch.from_numpy(Q.copy()).pin_memory()
O1 = torch.empty(n * 128, dtype=torch.uint8).pin_memory(); O2 = torch.empty_like(O1).pin_memory()
t0 = time.perf_counter(); pr.apply_into(O1, Pp, Qp, n); d1 = time.perf_counter() - t0
pr.set_devices($N)
pr.apply_into(O2, Pp, Qp, n)
best = 1e9
for _ in range(3):
    t0 = time.perf_counter(); pr.apply_into(O2, Pp, Qp, n); best = min(best, time.perf_counter() - t0)
line = {"probe": "single_process_fanout", "devices": $N, "n": n, "pairings_per_s": n / best,
        "one_device_pairings_per_s_incl_first_call": n / d1, "same_as_one_device": bool(torch.equal(O1, O2))}
print(json.dumps(line)); open("gpurun_out/r2_fanout_n$N.json", "w").write(json.dumps(line))
PY
