"""A/B probe: parity on the fixtures + device-resident throughput of one build of the library
(PBC_B200_LIB selects the .so).  usage: python tools/gpu_variant_probe.py f d [n]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbc_b200.pairing import Pairing
from pbc_b200.params import PARAMS
import bench

names = [a for a in sys.argv[1:] if not a.isdigit()]
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 1 << 17)
for wn in names:
    w = bench.WORKLOADS[wn]
    g = json.load(open("tests/golden/%s.json" % w["param"]))["pairing"]
    pr = Pairing(PARAMS[w["param"]])
    cat = lambda xs: b"".join(bytes.fromhex(x) for x in xs)
    ok = pr.apply(cat(g["P"]), cat(g["Q"]), len(g["e"])) == cat(g["e"])
    P, Q = bench.make_inputs(w, n)
    dP, dQ = torch.from_numpy(P.copy()).cuda(), torch.from_numpy(Q.copy()).cuda()
    dO = torch.empty(n * pr.gt_len, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    pr.set_stage_profiling(True)
    for _ in range(2):
        pr.apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        pr.apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"lib": os.path.basename(os.environ.get("PBC_B200_LIB", "default")), "workload": wn, "n": n,
                      "parity": ok, "ms": ms, "per_s": n / ms * 1e3, "stage_ms": pr.stage_times()}))
