"""Prototype (design aid) of the isomorphic tower the Type F kernels use internally:

    reference basis   F_q^2 = F_q[s]/(s^2 - beta),  F_q^12 = F_q^2[x]/(x^6 - xi),  xi = -alpha
    internal basis    K     = F_q[i]/(i^2 + 1),     F_q^12 = K[z]/(z^6 - xi'),     xi' = a' + b' i small

    phi2(a + b s) = a + (sigma b) i,   sigma^2 = -beta            (needs q = 3 mod 4)
    Phi(sum c_j x^j) = sum phi2(c_j) tau^j z^j,   tau^6 xi' = phi2(xi)

find_basis() is the algorithm engine.cu's init_type_f mirrors (tests compare the constants)."""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS


def find_basis(q, beta, xi, max_small=6, max_sylow=1 << 16):
    """returns dict(sigma, xi_small=(a', b'), tau) or None"""
    if q % 4 != 3:
        return None
    Fq = O.PrimeField(q)
    K = O.QuadExt(Fq, q - 1)
    sigma = pow((-beta) % q, (q + 1) // 4, q)
    if sigma * sigma % q != (-beta) % q:
        return None
    xi1 = (xi[0], sigma * xi[1] % q)
    N = q * q - 1
    e2 = e3 = 0
    m = N
    while m % 2 == 0:
        m //= 2; e2 += 1
    while m % 3 == 0:
        m //= 3; e3 += 1
    S = 2 ** e2 * 3 ** e3
    if S > max_sylow:
        return None
    k = next(k for k in range(1, 7) if (1 + k * m) % 6 == 0)
    t = (1 + k * m) // 6                       # 6 t = 1 + k m
    g = K.pow(xi1, m)                          # generates the subgroup of order S
    for total in range(1, 2 * max_small + 1):  # candidates by increasing a' + b', b' >= 1
        for b in range(1, total + 1):
            a = total - b
            if a > max_small or b > max_small:
                continue
            xs = (a, b)
            c = K.mul(xi1, K.inv(xs))
            if K.pow(c, N // 6) != K.one:
                continue
            w = K.pow(c, t)                    # w^6 = c (c^m)^k
            D = K.inv(K.pow(K.pow(c, m), k))
            h = K.one
            for _ in range(S):
                h6 = K.pow(h, 6)
                if h6 == D:
                    tau = K.mul(w, h)
                    assert K.mul(K.pow(tau, 6), xs) == xi1
                    return dict(sigma=sigma, xi_small=xs, tau=tau, xi1=xi1)
                h = K.mul(h, g)
    return None


def main():
    pr = O.pairing_from_param(PARAMS["f"])
    q, F2, F12 = pr.q, pr.Fq2, pr.Fq12
    beta, xi = F2.nqr, pr.negalpha
    B = find_basis(q, beta, xi)
    assert B is not None
    sigma, xs, tau = B["sigma"], B["xi_small"], B["tau"]
    print("q mod 4 =", q % 4, " xi' =", xs)
    K = O.QuadExt(pr.Fq, q - 1)
    F12n = O.PolyModExt(K, [K.neg(xs)] + [K.zero] * 5)      # z^6 = xi'
    phi2 = lambda c: (c[0], sigma * c[1] % q)
    taup = [K.pow(tau, j) for j in range(6)]
    Phi = lambda f: tuple(K.mul(phi2(c), taup[j]) for j, c in enumerate(f))
    sinv = pow(sigma, -1, q)
    tinv = [K.inv(t) for t in taup]
    PhiInv = lambda f: tuple((lambda d: (d[0], d[1] * sinv % q))(K.mul(c, tinv[j])) for j, c in enumerate(f))
    rnd = random.Random(3)
    for _ in range(4):
        x = tuple((rnd.randrange(q), rnd.randrange(q)) for _ in range(6))
        y = tuple((rnd.randrange(q), rnd.randrange(q)) for _ in range(6))
        assert Phi(F12.mul(x, y)) == F12n.mul(Phi(x), Phi(y))
        assert PhiInv(Phi(x)) == x
        assert Phi(F12.inv(x)) == F12n.inv(Phi(x))
    # Frobenius in the internal basis: conj + xi'^(j (q^k - 1)/6)
    for k in (1, 2, 3):
        gam = K.pow(xs, (q ** k - 1) // 6)
        x = tuple((rnd.randrange(q), rnd.randrange(q)) for _ in range(6))
        want = Phi(F12.pow(x, q ** k))
        conj = (lambda c: (c[0], (-c[1]) % q)) if k & 1 else (lambda c: c)
        got = tuple(K.mul(conj(c), K.pow(gam, j)) for j, c in enumerate(Phi(x)))
        assert got == want
    print("isomorphism, inverse and Frobenius maps check out; tau =", tau)


if __name__ == "__main__":
    main()
