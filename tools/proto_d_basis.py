"""Prototype (design aid) of the internal cubic basis of the Type D tower:
    reference   F_q^3 = F_q[x]/(x^3 + c2 x^2 + c1 x + c0)
    internal    F_q^3 = F_q[w]/(w^3 + p w + 1),        x = lam w - s,  s = c2/3,  lam^3 = R0
with P = c1 - c2^2/3, R0 = c0 - c1 c2/3 + 2 c2^3/27, p = P/lam^2 (needs q = 2 mod 3 for the unique
cube root lam = R0^((2q-1)/3)).  A product then needs 6 + 2 multiplications instead of 6 + 6:
    w^3 = -p w - 1,  w^4 = -p w^2 - w."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS


def find_basis(q, c0, c1, c2):
    if q % 3 != 2:
        return None
    i3 = pow(3, -1, q)
    s = c2 * i3 % q
    P = (c1 - c2 * c2 * i3) % q
    R0 = (c0 - c1 * c2 * i3 + 2 * pow(c2, 3, q) * pow(27, -1, q)) % q
    lam = pow(R0, (2 * q - 1) // 3, q)
    assert pow(lam, 3, q) == R0
    p = P * pow(lam * lam, -1, q) % q
    return dict(s=s, lam=lam, p=p)


def main():
    pr = O.pairing_from_param(PARAMS["d159"])
    q, F3 = pr.q, pr.Fq3
    c0, c1, c2 = F3.low
    B = find_basis(q, c0, c1, c2)
    s, lam, p = B["s"], B["lam"], B["p"]
    F3n = O.PolyModExt(pr.Fq, [1, p, 0])                  # w^3 + p w + 1
    li = pow(lam, -1, q)

    def to_int(a):                                         # a0 + a1 x + a2 x^2, x = lam w - s
        a0, a1, a2 = a
        return ((a0 - a1 * s + a2 * s * s) % q, lam * (a1 - 2 * s * a2) % q, lam * lam * a2 % q)

    def to_ref(b):
        b0, b1, b2 = b
        a2 = b2 * li * li % q
        a1 = (b1 * li + 2 * s * a2) % q
        return ((b0 + a1 * s - a2 * s * s) % q, a1, a2)

    rnd = random.Random(4)
    for _ in range(5):
        x = tuple(rnd.randrange(q) for _ in range(3)); y = tuple(rnd.randrange(q) for _ in range(3))
        assert to_int(F3.mul(x, y)) == F3n.mul(to_int(x), to_int(y))
        assert to_ref(to_int(x)) == x
    # Frobenius constants in the internal basis
    wq = F3n.pow((0, 1, 0), q)
    x = tuple(rnd.randrange(q) for _ in range(3))
    xi = to_int(x)
    fr = F3n.add(F3n.add((xi[0], 0, 0), F3n.scale(wq, xi[1])), F3n.scale(F3n.sqr(wq), xi[2]))
    assert fr == to_int(F3.pow(x, q))
    print("internal cubic w^3 + p w + 1: isomorphism, inverse and Frobenius check out; p =", p)


if __name__ == "__main__":
    main()
