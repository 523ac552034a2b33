"""Prototype of the Type A Miller loop in weight-(1,2) coordinates (x = X/Z, y = Y/Z^2), the
doubling of Costello-Lange-Naehrig for y^2 = x^3 + a x re-derived for a = 1:
  A = X^2, B = Y^2, C = Z^2;  X' = (A - C)^2,  Z' = 4 B,
  Y' = (2 (A + C)^2 - X') ((A - C + Y)^2 - B - X')
  line at phi(Q) (up to F_q^*):  Re = X (A - C) + (3A + C) Z Qx,   Im = ((Y + Z)^2 - B - C) Qy
5 M + 7 S per doubling (Jacobian: 8 M + 6 S).  Checked against the reference fixtures."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS
from tools.proto_a import finalexp_dev


def miller_w12(pr, P, Q):
    p = pr.q
    X, Y = P; Z = 1
    F0, F1 = 1, 0
    QX, QY = Q
    for i in range(pr.exp2):
        if i == pr.exp1:
            V1 = (X, Y if pr.sign1 > 0 else (-Y) % p, Z)
            G = (F0, F1 if pr.sign1 > 0 else (-F1) % p)
        T0 = (F0 + F1) % p; T1 = (F0 - F1) % p; F1 = 2 * F0 * F1 % p; F0 = T0 * T1 % p
        A = X * X % p; B = Y * Y % p; C = Z * Z % p
        T3 = (A - C) % p; T4 = (A + C) % p
        G3 = (2 * A + T4) % p
        L0 = (G3 * Z % p * QX + X * T3) % p
        L1 = ((Y + Z) ** 2 - B - C) % p * QY % p
        Zn = 4 * B % p
        Xn = T3 * T3 % p
        E = (2 * T4 * T4 - Xn) % p
        Fh = ((T3 + Y) ** 2 - B - Xn) % p
        X, Y, Z = Xn, E * Fh % p, Zn
        a = F0 * L0 % p; b = F1 * L1 % p; c = (F0 + F1) * (L0 + L1) % p
        F0 = (a - b) % p; F1 = (c - a - b) % p
    a = F0 * G[0] % p; b = F1 * G[1] % p; c = (F0 + F1) * (G[0] + G[1]) % p
    F0 = (a - b) % p; F1 = (c - a - b) % p
    X1, Y1, Z1 = V1
    la = (Y * Z1 * Z1 - Y1 * Z * Z) % p
    lb = Z * Z1 % p * (X1 * Z - X * Z1) % p
    lc = (X * Z % p * Y1 - Y * X1 % p * Z1) % p
    L0 = (lc - la * QX) % p; L1 = lb * QY % p
    a = F0 * L0 % p; b = F1 * L1 % p; c = (F0 + F1) * (L0 + L1) % p
    return ((a - b) % p, (c - a - b) % p)


if __name__ == "__main__":
    pr = O.pairing_from_param(PARAMS["a"])
    g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "a.json")))
    for Pb, Qb, e in zip(g["pairing"]["P"], g["pairing"]["Q"], g["pairing"]["e"]):
        P = pr.G1.from_bytes(bytes.fromhex(Pb)); Q = pr.G2.from_bytes(bytes.fromhex(Qb))
        assert pr.GT.to_bytes(finalexp_dev(pr, miller_w12(pr, P, Q))).hex() == e
    print("weight-(1,2) Miller loop == reference on", len(g["pairing"]["e"]), "vectors")
