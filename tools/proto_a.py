"""Prototype of the device algorithm for Type A (inversion-free Miller loop + shared-inversion
final exponentiation), on Python ints, checked against the oracle.  Design aid, not product."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

def miller_dev(pr, P, Q):
    p = pr.q
    X, Y = P; Z = 1; Z2 = 1
    F0, F1 = 1, 0
    QX, QY = Q
    for i in range(pr.exp2):
        if i == pr.exp1:
            V1 = (X, Y if pr.sign1 > 0 else (-Y) % p, Z)
            G = (F0, F1 if pr.sign1 > 0 else (-F1) % p)   # 1/f ~ conj(f) up to an Fq factor
        # f = f^2
        T0 = (F0 + F1) % p; T1 = (F0 - F1) % p; F1 = 2 * F0 * F1 % p; F0 = T0 * T1 % p
        T0 = X * X % p; T1 = Z2 * Z2 % p
        M = (3 * T0 + T1) % p
        Y2 = Y * Y % p
        S = 4 * X * Y2 % p
        T3 = M * Z2 % p
        T4 = T3 * QX % p
        T5 = (X * M - 2 * Y2) % p
        L0 = (T5 + T4) % p
        Zn = 2 * Y * Z % p
        L1 = Zn * Z2 % p * QY % p
        Z2 = Zn * Zn % p; Z = Zn
        Xn = (M * M - 2 * S) % p
        Y = (M * (S - Xn) - 8 * Y2 * Y2) % p
        X = Xn
        # f *= l
        a = F0 * L0 % p; b = F1 * L1 % p; c = (F0 + F1) * (L0 + L1) % p
        F0 = (a - b) % p; F1 = (c - a - b) % p
    # f *= f1
    a = F0 * G[0] % p; b = F1 * G[1] % p; c = (F0 + F1) * (G[0] + G[1]) % p
    F0 = (a - b) % p; F1 = (c - a - b) % p
    X1, Y1, Z1 = V1
    Z3 = Z * Z % p * Z % p; Z13 = Z1 * Z1 % p * Z1 % p
    la = (Y * Z13 - Y1 * Z3) % p
    lb = (X1 * Z1 % p * Z3 - X * Z % p * Z13) % p
    lc = (X * Z % p * Y1 - Y * X1 % p * Z1) % p
    L0 = (lc - la * QX) % p; L1 = lb * QY % p
    a = F0 * L0 % p; b = F1 * L1 % p; c = (F0 + F1) * (L0 + L1) % p
    return ((a - b) % p, (c - a - b) % p)

def finalexp_dev(pr, f):
    p = pr.q
    f0, f1 = f
    N = (f0 * f0 + f1 * f1) % p; W = f0 * f1 % p
    Dinv = pow(N * W % p, -1, p)          # the one inversion, batched across pairings on device
    Ninv = Dinv * W % p
    P_ = 2 * (f0 * f0 - f1 * f1) % p * Ninv % p
    v0, v1 = O._lucas_ladder(pr.Fq, 2, P_, pr.h)
    out0 = v0 * pow(2, -1, p) % p
    out1 = (2 * v1 - P_ * v0) % p * N % p * N % p * Dinv % p * pow(8, -1, p) % p
    return (out0, out1)

if __name__ == "__main__":
    pr = O.pairing_from_param(PARAMS["a"])
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "a.json")))
    for Pb, Qb, e in zip(g["pairing"]["P"], g["pairing"]["Q"], g["pairing"]["e"]):
        P = pr.G1.from_bytes(bytes.fromhex(Pb)); Q = pr.G2.from_bytes(bytes.fromhex(Qb))
        got = pr.GT.to_bytes(finalexp_dev(pr, miller_dev(pr, P, Q))).hex()
        assert got == e
    print("type A device algorithm == reference on", len(g["pairing"]["e"]), "vectors")
