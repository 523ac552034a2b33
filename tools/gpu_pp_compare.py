"""plain pairings vs the pairing_pp_t route (pbc_b200_pp_init once + pbc_b200_pp_apply_device), device-resident,
per parameter set: JSON lines {type, n, plain_per_s, pp_per_s, ratio, same_bytes}"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]] + sys.argv[1:]
import torch  # noqa: E402
import bench  # noqa: E402
from pbc_b200.pairing import Pairing  # noqa: E402
from pbc_b200.params import PARAMS  # noqa: E402

names = [a for a in sys.argv[1:] if not a.isdigit()] or ["a", "f", "d", "g"]
n = next((int(a) for a in sys.argv[1:] if a.isdigit()), 1 << 17)
for wn in names:
    w = bench.WORKLOADS[wn]
    pr = Pairing(PARAMS[w["param"]])
    P, Q = bench.make_inputs(w, n)
    g1 = pr.g1_len
    P1 = P[:g1].copy()
    Pn = torch.from_numpy(P1).repeat(n).cuda()            # plain path: the same P in every pair
    dQ = torch.from_numpy(Q.copy()).cuda()
    dO = torch.empty(n * pr.gt_len, dtype=torch.uint8, device="cuda")
    dO2 = torch.empty_like(dO)
    st = torch.cuda.current_stream()
    h = pr.pp_init(P1.tobytes())

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(3):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3

    t_plain = timed(lambda: pr.apply_device(dO.data_ptr(), Pn.data_ptr(), dQ.data_ptr(), n, st.cuda_stream))
    t_pp = timed(lambda: h.apply_device(dO2.data_ptr(), dQ.data_ptr(), n, st.cuda_stream))
    print(json.dumps({"type": w["param"], "n": n, "plain_per_s": n / t_plain * 1e3, "pp_per_s": n / t_pp * 1e3,
                      "ratio": t_plain / t_pp, "same_bytes": bool(torch.equal(dO, dO2))}), flush=True)
    h.clear()
