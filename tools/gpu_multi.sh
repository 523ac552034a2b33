#!/bin/bash
# N-GPU scaling check: one rank per GPU under torchrun, as the driver launches it
mkdir -p gpurun_out
N=${NGPU:-2}
for w in a prod16; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 --workload $w > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err; echo "bench $w N=$N rc=$?"
tail -2 gpurun_out/bench_${w}_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_n$N.json')); print('$w', d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'])"
done
timeout 300 python bench.py --impl reference --gpus $N --steps 1 --warmup 3 --ref-seconds 1 | cut -c1-300
# single-process fan-out over N devices through the C ABI (pbc_b200_set_devices)
timeout 600 python - <<PY
import json, time, torch
import bench
from pbc_b200.pairing import Pairing
from pbc_b200.params import PARAMS
w = bench.WORKLOADS["a"]; n = 1 << 20
P, Q = bench.make_inputs(w, n)
pr = Pairing(PARAMS["a"])
Pp, Qp = torch.from_numpy(P.copy()).pin_memory(), torch.from_numpy(Q.copy()).pin_memory()
O1 = torch.empty(n * 128, dtype=torch.uint8).pin_memory(); O2 = torch.empty_like(O1).pin_memory()
pr.apply_into(O1, Pp, Qp, n)
pr.set_devices($N)
pr.apply_into(O2, Pp, Qp, n)
t0 = time.perf_counter(); pr.apply_into(O2, Pp, Qp, n); dt = time.perf_counter() - t0
print(json.dumps({"probe": "single_process_fanout", "devices": $N, "n": n, "pairings_per_s": n / dt, "same_as_one_device": bool((O1 == O2).all())}))
PY
