"""Generate tests/golden/edge.json: the edge cases SURVEY 8(c) names, expected bytes from the compiled,
unmodified reference (oracle/_ref).  Run HERE:  make -C oracle && python tests/golden/make_edge_golden.py

Cases per parameter set (label -> P, Q, e), all in the reference wire format, hex:
  type a (G1 = G2):  P = Q, P = -Q, P / Q / both on the curve but OUTSIDE the order-r subgroup, products and
                     fixed-argument pairings that contain such points;
  types f, d159, g149:  -P, -Q, P on E(F_q) outside the order-r subgroup (d159, g149: cofactor > 1),
                     coordinates of Q written unreduced (x + q where it still fits the 20 / 19 bytes).
The 2-torsion input (0, 0) of type a is NOT taken from the reference (it inverts zero and returns a
by-product of that); include/pbc_b200.h documents what this library does instead, and the test checks that.
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402
from pbc_b200.params import PARAMS  # noqa: E402
from pbc_b200 import synth  # noqa: E402

SEED = 20260923


def sqrt_mod(a, q):
    """Tonelli-Shanks; None if a is not a square"""
    a %= q
    if a == 0:
        return 0
    if pow(a, (q - 1) // 2, q) != 1:
        return None
    if q % 4 == 3:
        return pow(a, (q + 1) // 4, q)
    s, t = q - 1, 0
    while s % 2 == 0:
        s //= 2
        t += 1
    z = 2
    while pow(z, (q - 1) // 2, q) != q - 1:
        z += 1
    c, x, b, m = pow(z, s, q), pow(a, (s + 1) // 2, q), pow(a, s, q), t
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % q
            i += 1
        g = pow(c, 1 << (m - i - 1), q)
        x, c, b, m = x * g % q, g * g % q, b * g * g % q, i
    return x


def curve_points(q, a, b, count, rnd, wb):
    """random points of y^2 = x^3 + a x + b over F_q (no cofactor multiplication)"""
    out = []
    while len(out) < count:
        x = rnd.randrange(q)
        y = sqrt_mod(x * x * x + a * x + b, q)
        if y is None or y == 0:
            continue
        if rnd.random() < 0.5:
            y = q - y
        out.append(x.to_bytes(wb, "big") + y.to_bytes(wb, "big"))
    return out


def neg_point(pt, q, wb, ncoord):
    """-(x, y): negate every coordinate of y (element of F_q^ncoord)"""
    xs, ys = pt[:wb * ncoord], pt[wb * ncoord:]
    ny = b"".join(((q - int.from_bytes(ys[i * wb:(i + 1) * wb], "big")) % q).to_bytes(wb, "big") for i in range(ncoord))
    return xs + ny


def main():
    doc = {"source": "oracle/_ref/libpbcref.so (unmodified reference), tests/golden/make_edge_golden.py, seed %d" % SEED}
    for name in ("a", "f", "d159", "g149"):
        prm = synth.parse_param(PARAMS[name])
        q = prm["q"]
        wb = (q.bit_length() + 7) // 8
        rp = R.RefPairing(PARAMS[name])
        R.RefPairing.seed(SEED)
        rnd = random.Random(SEED + len(name))
        g1, g2 = rp.g1_len, rp.g2_len
        n2 = g2 // (2 * wb)                       # coordinates of G2 live in F_q^n2
        P = rp.random(R.G1, 4)
        Q = rp.random(R.G2, 4)
        P0, P1, Q0, Q1 = P[:g1], P[g1:2 * g1], Q[:g2], Q[g2:2 * g2]
        cases = []

        def add(label, p, qq):
            cases.append({"label": label, "P": p.hex(), "Q": qq.hex(), "e": rp.pairing(p, qq, 1).hex()})

        add("plain", P0, Q0)
        add("-P", neg_point(P0, q, wb, 1), Q0)
        add("-Q", P0, neg_point(Q0, q, wb, n2))
        add("-P,-Q", neg_point(P0, q, wb, 1), neg_point(Q0, q, wb, n2))
        if name == "a":
            add("P = Q", P0, P0)
            add("P = -Q", P0, neg_point(P0, q, wb, 1))
            pts = curve_points(q, 1, 0, 3, rnd, wb)
            add("P outside the subgroup", pts[0], Q0)
            add("Q outside the subgroup", P0, pts[1])
            add("both outside the subgroup", pts[0], pts[1])
            add("P = Q outside the subgroup", pts[2], pts[2])
            extra = pts
        else:
            acoef = prm.get("a", 0)
            pts = curve_points(q, acoef, prm["b"], 2, rnd, wb)
            add("P on E(F_q), no cofactor multiplication", pts[0], Q0)
            add("another such P", pts[1], Q1)
            extra = pts
            # unreduced coordinates: x + q still fits wb bytes for some coordinate?
            qq = bytearray(Q0)
            done = False
            for ci in range(2 * n2):
                v = int.from_bytes(Q0[ci * wb:(ci + 1) * wb], "big")
                if v + q < (1 << (8 * wb)):
                    qq[ci * wb:(ci + 1) * wb] = (v + q).to_bytes(wb, "big")
                    done = True
            if done:
                add("Q with coordinates written as x + q", P0, bytes(qq))
            pp_ = bytearray(P0)
            v = int.from_bytes(P0[:wb], "big")
            if v + q < (1 << (8 * wb)):
                pp_[:wb] = (v + q).to_bytes(wb, "big")
                add("P with x written as x + q", bytes(pp_), Q0)
        # a product and a fixed-argument batch over the special points
        k = 3
        PP = P0 + extra[0] + neg_point(P1, q, wb, 1)
        QQ = Q0 + (extra[1] if name == "a" else Q1) + Q1
        prod = {"k": k, "P": PP.hex(), "Q": QQ.hex(), "e": rp.prod_pairing(PP, QQ, k, 1).hex()}
        ppQ = Q0 + neg_point(Q1, q, wb, n2) + (extra[1] if name == "a" else Q1)
        pp = {"P": extra[0].hex(), "Q": ppQ.hex(), "e": rp.pp_pairing(extra[0], ppQ, 3).hex()}
        doc[name] = {"cases": cases, "prod": prod, "pp": pp}
    path = os.path.join(ROOT, "tests", "golden", "edge.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", path, {k: len(v["cases"]) for k, v in doc.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
