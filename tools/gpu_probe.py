"""First-light GPU probe: integer-pipe peak, F_p multiplier throughput vs occupancy, Type A
pairing throughput (device-resident).  Prints one JSON object per line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbc_b200.pairing import Pairing, bench_imad, kernel_launches
from pbc_b200.params import PARAMS

def main():
    pr = Pairing(PARAMS["a"])
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "sms": sm}))
    for blocks_per_sm, threads in ((4, 256), (8, 256), (2, 1024)):
        iters = 4000
        ms = bench_imad(sm * blocks_per_sm, threads, iters, 3)
        n = sm * blocks_per_sm * threads * 32 * iters
        print(json.dumps({"probe": "imad_wide_peak", "blocks_per_sm": blocks_per_sm, "threads": threads,
                          "ms": ms, "imad_per_s": n / (ms * 1e-3)}))
    for mode in (0, 2, 3, 1):
        for bps in (1, 2, 3, 4, 8):
            iters = 2000
            ms = pr.bench_fpmul(mode, sm * bps, iters, 3)
            muls = sm * bps * 128 * iters
            print(json.dumps({"probe": "fpmul512", "mode": {0: "regs_operand_scan", 1: "slots_kernel_mult", 2: "regs_product_scan", 3: "regs_product_scan_sqr"}[mode], "blocks_per_sm": bps,
                              "warps_per_sm": bps * 4, "ms": ms, "mulmod_per_s": muls / (ms * 1e-3),
                              "imad_equiv_per_s": 528 * muls / (ms * 1e-3)}))
    g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "a.json")))["pairing"]
    P = b"".join(bytes.fromhex(x) for x in g["P"]); Q = b"".join(bytes.fromhex(x) for x in g["Q"])
    m = len(g["e"])
    for n in (sm * 256, sm * 256 * 4, 1 << 18):
        reps = n // m + 1
        dP = torch.frombuffer(bytearray((P * reps)[:n * 128]), dtype=torch.uint8).cuda()
        dQ = torch.frombuffer(bytearray((Q * reps)[:n * 128]), dtype=torch.uint8).cuda()
        out = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream()
        pr.apply_device(out.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        pr.apply_device(out.data_ptr(), dP.data_ptr(), dQ.data_ptr(), n, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"probe": "pairing_a_device", "n": n, "ms": ms, "pairings_per_s": n / (ms * 1e-3)}))
    print(json.dumps({"launches": kernel_launches()}))

if __name__ == "__main__":
    main()
