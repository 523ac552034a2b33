"""Prototype (design aid): Granger-Scott squaring in the cyclotomic subgroup for the Type F tower.
F_q^12 = K[z]/(z^6 - xi) seen as a cubic extension of F_q^4 = K[s]/(s^2 - xi), s = z^3:
    f = u0 + u1 z + u2 z^2,  u0 = (c0, c3), u1 = (c1, c4), u2 = (c2, c5)
    f^2 = (3 u0^2 - 2 conj(u0)) + (3 s u2^2 + 2 conj(u1)) z + (3 u1^2 - 2 conj(u2)) z^2
valid for f of norm 1 over F_q^6 with f^(q^4 - q^2 + 1) = 1, i.e. after the easy part of the final
exponentiation.  Works in either basis (xi = reference -alpha, or the internal xi')."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS


def main():
    pr = O.pairing_from_param(PARAMS["f"])
    q, K, F12 = pr.q, pr.Fq2, pr.Fq12
    xi = pr.negalpha

    def sq4(a, b):               # (a + b s)^2, s^2 = xi
        t = K.mul(a, b)
        return (K.add(K.sqr(a), K.mul(xi, K.sqr(b))), K.add(t, t))

    def cyc_sqr(f):
        c = f
        u0, u1, u2 = (c[0], c[3]), (c[1], c[4]), (c[2], c[5])
        A, B, C = sq4(*u0), sq4(*u1), sq4(*u2)
        sC = (K.mul(xi, C[1]), C[0])            # s * (C0 + C1 s) = xi C1 + C0 s
        three = lambda x: K.add(K.add(x, x), x)
        two = lambda x: K.add(x, x)
        n0 = (K.sub(three(A[0]), two(u0[0])), K.add(three(A[1]), two(u0[1])))     # 3A - 2 conj(u0)
        n1 = (K.add(three(sC[0]), two(u1[0])), K.sub(three(sC[1]), two(u1[1])))   # 3 s C + 2 conj(u1)
        n2 = (K.sub(three(B[0]), two(u2[0])), K.add(three(B[1]), two(u2[1])))     # 3B - 2 conj(u2)
        return (n0[0], n1[0], n2[0], n0[1], n1[1], n2[1])

    rnd = random.Random(9)
    for _ in range(5):
        f = tuple((rnd.randrange(q), rnd.randrange(q)) for _ in range(6))
        conj = tuple(c if i % 2 == 0 else K.neg(c) for i, c in enumerate(f))
        g = F12.mul(conj, F12.inv(f))                       # ^(q^6 - 1)
        g = F12.mul(F12.pow(g, q * q), g)                   # ^(q^2 + 1): now cyclotomic
        assert cyc_sqr(g) == F12.sqr(g)
    print("Granger-Scott cyclotomic squaring == generic squaring on cyclotomic elements")


if __name__ == "__main__":
    main()
