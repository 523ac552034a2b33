"""Battery run against the CPU simulator of the product library (PBC_B200_LIB must point at it):
every entry point of the C ABI for every parameter set, compared with the reference fixtures.
Prints one JSON object {check: bool}.  TEST INFRASTRUCTURE (tests/test_kernels_on_cpu_sim.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pbc_b200.pairing import Pairing  # noqa: E402
from pbc_b200.params import PARAMS  # noqa: E402

cat = lambda xs: b"".join(bytes.fromhex(x) for x in xs)  # noqa: E731
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
res = {}


def golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name + ".json")) as f:
        return json.load(f)


if mode == "all":
    for name in ("a", "f", "d159", "g149"):
        g = golden(name)
        pr = Pairing(PARAMS[name])
        n = 2
        res[name + ".pairing"] = pr.apply(cat(g["pairing"]["P"][:n]), cat(g["pairing"]["Q"][:n]), n) == cat(g["pairing"]["e"][:n])
        k, no = g["prod"]["k"], len(g["prod"]["e"])
        res[name + ".prod"] = pr.prod_apply(cat(g["prod"]["P"]), cat(g["prod"]["Q"]), k, no) == cat(g["prod"]["e"])
        res[name + ".pp"] = pr.pp_apply(bytes.fromhex(g["pp"]["P"]), cat(g["pairing"]["Q"][:4]), 4) == cat(g["pp"]["e"])
        res[name + ".g1_pow"] = pr.g1_pow_zn(cat(g["pairing"]["P"][:4]), cat(g["pow"]["a"]), 4) == cat(g["pow"]["Pa"])
        res[name + ".g2_pow"] = pr.g2_pow_zn(cat(g["pairing"]["Q"][:4]), cat(g["pow"]["a"]), 4) == cat(g["pow"]["Qa"])
        blk = g["hash"]["32"]
        res[name + ".from_hash"] = pr.g1_from_hash(cat(blk["data"]), 32, len(blk["data"])) == cat(blk["G1"])
        nc = len(g["compressed"]["G1"])
        res[name + ".decompress"] = pr.g1_decompress(cat(g["compressed"]["G1"]), nc) == cat(g["pairing"]["P"][:nc])
        ident = bytes.fromhex(g["offcurve"]["identity"])
        res[name + ".offcurve"] = pr.apply(bytes.fromhex(g["offcurve"]["badP"]), cat(g["pairing"]["Q"][:1]), 1) == ident

# type A1: the small parameter set in full, one pairing of the 1033-bit one
g = golden("a1_small")
pr = Pairing(g["param_text"])
n = len(g["pairing"]["e"])
res["a1_small.pairing"] = pr.apply(cat(g["pairing"]["P"]), cat(g["pairing"]["Q"]), n) == cat(g["pairing"]["e"])
k, no = g["prod"]["k"], len(g["prod"]["e"])
res["a1_small.prod"] = pr.prod_apply(cat(g["prod"]["P"][:k * no]), cat(g["prod"]["Q"][:k * no]), k, no) == cat(g["prod"]["e"])
res["a1_small.pp"] = pr.pp_apply(bytes.fromhex(g["pp"]["P"]), cat(g["pairing"]["Q"][:4]), 4) == cat(g["pp"]["e"])
res["a1_small.offcurve"] = pr.apply(bytes.fromhex(g["offcurve"]["badP"]), cat(g["pairing"]["Q"][:1]), 1) == bytes.fromhex(g["offcurve"]["identity"])
g = golden("a1")
pr = Pairing(PARAMS["a1"])
res["a1.pairing"] = pr.apply(cat(g["pairing"]["P"][:1]), cat(g["pairing"]["Q"][:1]), 1) == cat(g["pairing"]["e"][:1])
res["a1.pp"] = pr.pp_apply(bytes.fromhex(g["pp"]["P"]), cat(g["pairing"]["Q"][:1]), 1) == cat(g["pp"]["e"][:1])
print(json.dumps(res))
