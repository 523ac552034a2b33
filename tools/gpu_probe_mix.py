"""Occupancy sensitivity of the slot machine with a Miller-like multiply/add mix."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pbc_b200.pairing import Pairing
from pbc_b200.params import PARAMS
pr = Pairing(PARAMS["a"])
sm = torch.cuda.get_device_properties(0).multi_processor_count
for mode, name in ((1, "mul"), (4, "mul+add calls"), (5, "fused mul+add")):
    for bps in (1, 2, 3, 4, 6):
        iters = 2000
        ms = pr.bench_fpmul(mode, sm * bps, iters, 3)
        muls = sm * bps * 128 * iters
        print(json.dumps({"probe": "slots512", "mix": name, "warps_per_sm": bps * 4, "ms": ms,
                          "mulmod_per_s": muls / (ms * 1e-3), "frac_of_17.05G": muls / (ms * 1e-3) / 17.05e9}))
