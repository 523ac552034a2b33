"""Prototype (Python integers) of the inversion-free Miller loop the F/D CUDA kernels run:
Jacobian multiples of P, line coefficients scaled by elements of F_q^*, checked against the
oracle AFTER the final exponentiation (the scale factors must vanish)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS


def miller_proj(Fq, A, r, P, mul_line, one, sqr):
    q = Fq.p
    xP, yP = P
    X, Y, Z = xP, yP, 1
    v = one
    m = r.bit_length() - 2
    while True:
        Z2 = Z * Z % q
        M = (3 * X * X + A * Z2 * Z2) % q
        Y2 = Y * Y % q
        Zn = 2 * Y * Z % q
        a = (-M * Z2) % q
        b = Zn * Z2 % q
        c = (M * X - 2 * Y2) % q
        v = mul_line(v, a, b, c)
        if m == 0:
            break
        # V = 2V
        S = 4 * X * Y2 % q
        Xn = (M * M - 2 * S) % q
        Yn = (M * (S - Xn) - 8 * Y2 * Y2) % q
        X, Y, Z = Xn, Yn, Zn
        if (r >> m) & 1:
            Z2 = Z * Z % q
            Z3 = Z2 * Z % q
            H = (xP * Z2 - X) % q
            Rr = (yP * Z3 - Y) % q
            a = (-Rr) % q                    # Y - yP Z^3
            b = H * Z % q                    # (xP Z^2 - X) Z  == Z3 of the sum
            c = (yP * Z * X - xP * Y) % q
            v = mul_line(v, a, b, c)
            H2 = H * H % q
            H3 = H2 * H % q
            XH2 = X * H2 % q
            Xn = (Rr * Rr - H3 - 2 * XH2) % q
            Yn = (Rr * (XH2 - Xn) - Y * H3) % q
            X, Y, Z = Xn, Yn, b
        m -= 1
        v = sqr(v)
    return v


def main():
    for name in ("f", "d159"):
        g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".json")))
        pr = O.pairing_from_param(PARAMS[name])
        for i in range(3):
            P = pr.G1.from_bytes(bytes.fromhex(g["pairing"]["P"][i]))
            Q = pr.G2.from_bytes(bytes.fromhex(g["pairing"]["Q"][i]))
            if name == "f":
                Qx = pr.Fq2.mul(Q[0], pr.negalphainv); Qy = pr.Fq2.mul(Q[1], pr.negalphainv)
                f = miller_proj(pr.Fq, pr.E.a, pr.r, P, pr._mul_line(Qx, Qy), pr.Fq12.one, pr.Fq12.sqr)
                out = pr.GT.to_bytes(pr.final_exp(f))
            else:
                Qx = pr.Fq3.mul(Q[0], pr.nqrinv); Qy = pr.Fq3.mul(Q[1], pr.nqrinv2)
                f = miller_proj(pr.Fq, pr.E.a, pr.r, P, pr._mul_line(Qx, Qy), pr.Fq6.one, pr.Fq6.sqr)
                out = pr.GT.to_bytes(pr.tatepower(f))
            print(name, i, out.hex() == g["pairing"]["e"][i])


if __name__ == "__main__":
    main()
