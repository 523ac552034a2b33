"""pp (fixed first argument) path at growing n: device and host entry points, timed, checked against
plain pairings with the same P."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pbc_b200.pairing import Pairing
from pbc_b200.params import PARAMS
w = bench.WORKLOADS["a"]
for n in (1 << 12, 1 << 16, 1 << 18, (1 << 18) + 1000):
    P, Q = bench.make_inputs(w, n)
    pr = Pairing(PARAMS["a"])
    P0 = P[:128].tobytes()
    t0 = time.perf_counter(); got = pr.pp_apply(P0, Q.tobytes(), n); t_host = time.perf_counter() - t0
    t0 = time.perf_counter(); want = pr.apply(P0 * n, Q.tobytes(), n); t_plain = time.perf_counter() - t0
    print(json.dumps({"n": n, "pp_host_s": t_host, "plain_host_s": t_plain, "same": got == want}), flush=True)
    dP = torch.frombuffer(bytearray(P0), dtype=torch.uint8).cuda()
    dQ = torch.from_numpy(Q.copy()).cuda()
    dO = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    pr.pp_apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    pr.pp_apply_device(dO.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    print(json.dumps({"n": n, "pp_device_ms": e0.elapsed_time(e1), "per_s": n / e0.elapsed_time(e1) * 1e3,
                      "same": bytes(dO.cpu().numpy().tobytes()) == want}), flush=True)
