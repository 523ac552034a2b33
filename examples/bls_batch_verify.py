#!/usr/bin/env python
"""BLS signature verification, batched, entirely on the GPU (example/bls.c of the reference, scaled
up): every step is one of the batched entry points of include/pbc_b200.h.

    keygen   pk = sk * g2                                   pbc_b200_g2_pow_zn
    sign     sigma_i = sk * H(m_i)                          pbc_b200_g1_from_hash, pbc_b200_g1_pow_zn
    verify   e(sigma_i, g2) * e(-H(m_i), pk) == 1           pbc_b200_g1_from_hash, pbc_b200_prod_pairings_apply (k = 2)

usage: python examples/bls_batch_verify.py [a|a1|f|d159|g149] [n]
"""
import hashlib
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbc_b200.pairing import Pairing  # noqa: E402
from pbc_b200.params import PARAMS    # noqa: E402
from pbc_b200 import synth            # noqa: E402


def negate_points(points: bytes, g1_len: int, q: int) -> bytes:
    """-(x, y) = (x, q - y) on wire bytes (host side: a subtraction per point)"""
    half = g1_len // 2
    out = bytearray(points)
    for i in range(len(points) // g1_len):
        y = int.from_bytes(points[i * g1_len + half:(i + 1) * g1_len], "big")
        out[i * g1_len + half:(i + 1) * g1_len] = ((q - y) % q).to_bytes(half, "big")
    return bytes(out)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "a"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 14
    prm = synth.parse_param(PARAMS[name])
    pr = Pairing(PARAMS[name])
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    g2 = bytes.fromhex(g["pairing"]["Q"][0])                    # a fixed generator of G2
    rnd = random.Random(2026)
    order = prm["n"] if name == "a1" else prm["r"]              # a1.param names the group order n, the field p
    field = prm["p"] if name == "a1" else prm["q"]
    sk = rnd.randrange(1, order).to_bytes(pr.zr_len, "big")
    pk = pr.g2_pow_zn(g2, sk, 1)
    msgs = b"".join(hashlib.sha256(b"message %d" % i).digest() for i in range(n))

    t0 = time.perf_counter()
    H = pr.g1_from_hash(msgs, 32, n)
    sig = pr.g1_pow_zn(H, sk * n, n)
    t_sign = time.perf_counter() - t0

    # one forged signature to show the check bites
    bad = n // 2
    forged = bytearray(sig)
    forged[bad * pr.g1_len:(bad + 1) * pr.g1_len] = sig[:pr.g1_len]
    sig_in = bytes(forged)

    t0 = time.perf_counter()
    H2 = pr.g1_from_hash(msgs, 32, n)
    negH = negate_points(H2, pr.g1_len, field)
    in1 = b"".join(sig_in[i * pr.g1_len:(i + 1) * pr.g1_len] + negH[i * pr.g1_len:(i + 1) * pr.g1_len] for i in range(n))
    in2 = (g2 + pk) * n
    res = pr.prod_apply(in1, in2, 2, n)
    t_verify = time.perf_counter() - t0
    one = pr.gt_pow_zn(res[:pr.gt_len], bytes(pr.zr_len), 1)           # x^0: the GT identity in wire form
    ok = [res[i * pr.gt_len:(i + 1) * pr.gt_len] == one for i in range(n)]
    assert ok.count(False) == 1 and not ok[bad], "exactly the forged signature must fail"
    print(json.dumps({"example": "bls_batch_verify", "type": name, "n": n,
                      "sign_per_s": n / t_sign, "verify_per_s": n / t_verify,
                      "rejected": [i for i, v in enumerate(ok) if not v],
                      "note": "wall clock incl. host-side byte shuffling in Python; pairing work = 2 Miller loops + 1 final exponentiation per signature"}))


if __name__ == "__main__":
    main()
