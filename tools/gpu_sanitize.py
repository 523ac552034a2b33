"""One small batch of every hot path per parameter set, for compute-sanitizer (memcheck / synccheck / racecheck):
  compute-sanitizer --tool memcheck python tools/gpu_sanitize.py
Outputs are compared with the reference fixtures so a run that "passes" the tool also computed the right bytes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbc_b200.pairing import Pairing  # noqa: E402
from pbc_b200.params import PARAMS  # noqa: E402

cat = lambda xs: b"".join(bytes.fromhex(x) for x in xs)
names = sys.argv[1:] or ["a", "f", "d159", "g149", "a1"]
for name in names:
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    pr = Pairing(PARAMS[name])
    n = len(g["pairing"]["e"]) if name != "a1" else 3
    # ragged sizes on purpose: partial last block, partial last warp
    ok = pr.apply(cat(g["pairing"]["P"][:n]), cat(g["pairing"]["Q"][:n]), n) == cat(g["pairing"]["e"][:n])
    k = g["prod"]["k"]
    no = len(g["prod"]["e"]) if name != "a1" else 1
    ok &= pr.prod_apply(cat(g["prod"]["P"][:k * no]), cat(g["prod"]["Q"][:k * no]), k, no) == cat(g["prod"]["e"][:no])
    m = 3 if name != "a1" else 2
    ok &= pr.pp_apply(bytes.fromhex(g["pp"]["P"]), cat(g["pairing"]["Q"][:m]), m) == cat(g["pp"]["e"][:m])
    h = pr.pp_init(bytes.fromhex(g["pp"]["P"]))
    ok &= h.apply(cat(g["pairing"]["Q"][:m]), m) == cat(g["pp"]["e"][:m])
    h.clear()
    if name != "a1":
        ok &= pr.g1_pow_zn(cat(g["pairing"]["P"][:4]), cat(g["pow"]["a"]), 4) == cat(g["pow"]["Pa"])
        ok &= pr.gt_pow_zn(cat(g["pairing"]["e"][:4]), cat(g["pow"]["a"]), 4) == cat(g["pow"]["e_Pa_Q"])
    E = cat(g["pairing"]["e"][:2])
    ok &= pr.gt_cmp(E, E, 2) == b"\0\0"
    ok &= len(pr.gt_mul(E, E, 2)) == 2 * pr.gt_len
    print("sanitize batch %-5s %s" % (name, "bit-exact" if ok else "MISMATCH"), flush=True)
    pr.clear()
    if not ok:
        sys.exit(1)
print("sanitize batch: all parameter sets OK")
