"""Synthetic, seeded input batches in the reference wire format (bench.py, large-batch tests).

Host-side input synthesis only -- plain Python integers, no pairing arithmetic and no use of the
oracle.  A batch of n pairs is drawn from a g x g grid of distinct subgroup points:
P_i = P_0 + i*G (i < g), Q_j = Q_0 + j*H (j < g), pair k = (P[k mod g], Q[(k div g + 7 k) mod g]),
so every pair of a 2^20 batch is distinct for g = 4096 (SURVEY 8d config 2 builds its inputs the
same way: one affine addition per point instead of 2^21 scalar multiplications).
"""
from __future__ import annotations
import random


def _inv(a, q):
    return pow(a, -1, q)


class _Curve:
    """y^2 = x^3 + a x + b over F_q, affine, odd q."""

    def __init__(self, q, a, b):
        self.q, self.a, self.b = q, a, b

    def add(self, P, Q):
        q = self.q
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % q == 0:
                return None
            lam = (3 * x1 * x1 + self.a) * _inv(2 * y1, q) % q
        else:
            lam = (y2 - y1) * _inv(x2 - x1, q) % q
        x3 = (lam * lam - x1 - x2) % q
        return (x3, (lam * (x1 - x3) - y1) % q)

    def mul(self, k, P):
        R = None
        for bit in bin(k)[2:]:
            R = self.add(R, R)
            if bit == "1":
                R = self.add(R, P)
        return R

    def lift_x(self, rnd):
        """random point; needs q = 3 mod 4 for the square root"""
        q = self.q
        assert q % 4 == 3
        while True:
            x = rnd.randrange(q)
            rhs = (x * x * x + self.a * x + self.b) % q
            y = pow(rhs, (q + 1) // 4, q)
            if y * y % q == rhs:
                return (x, y if rnd.random() < 0.5 else (q - y) % q)


def type_a_points(param: dict, g: int, seed: int):
    """2 x g distinct points of the order-r subgroup of y^2 = x^3 + x (Type A, G1 = G2; Type A1
    with its keys p, l in place of q, h)."""
    q, h = (param["p"], param["l"]) if param.get("type") == "a1" else (param["q"], param["h"])
    wb = (q.bit_length() + 7) // 8
    E = _Curve(q, 1, 0)
    rnd = random.Random(seed)

    def subgroup_point():
        while True:
            P = E.mul(h, E.lift_x(rnd))
            if P is not None:
                return P

    def walk(start, step):
        out, cur = [], start
        for _ in range(g):
            out.append(cur)
            cur = E.add(cur, step)
        return out

    Ps = walk(subgroup_point(), subgroup_point())
    Qs = walk(subgroup_point(), subgroup_point())
    enc = lambda pt: pt[0].to_bytes(wb, "big") + pt[1].to_bytes(wb, "big")
    return [enc(p) for p in Ps], [enc(p) for p in Qs]


def parse_param(text: str) -> dict:
    out = {}
    for line in text.splitlines():
        t = line.split("#", 1)[0].split()
        if len(t) >= 2:
            out[t[0]] = t[1] if t[0] == "type" else int(t[1])
    return out


def pair_indices(n: int, g: int, offset: int = 0):
    """(i, j) grid coordinates of pairs offset .. offset+n-1"""
    for k in range(offset, offset + n):
        yield k % g, (k // g + 7 * k) % g


def build_batch(Pb, Qb, n: int, offset: int = 0):
    """numpy uint8 arrays (n*len(P), n*len(Q)) for pairs offset .. offset+n-1"""
    import numpy as np
    g = len(Pb)
    lp, lq = len(Pb[0]), len(Qb[0])
    Pa = np.frombuffer(b"".join(Pb), dtype=np.uint8).reshape(g, lp)
    Qa = np.frombuffer(b"".join(Qb), dtype=np.uint8).reshape(g, lq)
    k = np.arange(offset, offset + n, dtype=np.int64)
    return Pa[k % g].reshape(-1), Qa[(k // g + 7 * k) % g].reshape(-1)


# ---------------------------------------------------------------------------------------------
# Types F and D: grids of G1 / G2 points grown from seed points by repeated affine addition.
# G2 lives on a twist over F_q^2 (F) or F_q^3 (D); neither field has a cheap square root here
# (q = 1 mod 4 for d159), so the walks start from points the caller supplies (bench.py passes
# the committed reference fixtures): P_i = P_0 + i G stays in the subgroup of its seeds.
# ---------------------------------------------------------------------------------------------
class _ExtField:
    """F_q[x]/(x^n + m[n-1] x^(n-1) + ... + m[0]); elements are tuples of n residues."""

    def __init__(self, q, low):
        self.q, self.low, self.n = q, tuple(c % q for c in low), len(low)
        self.zero = (0,) * self.n
        self.one = (1,) + (0,) * (self.n - 1)

    def add(self, a, b):
        return tuple((x + y) % self.q for x, y in zip(a, b))

    def sub(self, a, b):
        return tuple((x - y) % self.q for x, y in zip(a, b))

    def mul(self, a, b):
        q, n = self.q, self.n
        d = [0] * (2 * n - 1)
        for i in range(n):
            for j in range(n):
                d[i + j] = (d[i + j] + a[i] * b[j]) % q
        for k in range(2 * n - 2, n - 1, -1):
            t, d[k] = d[k], 0
            for i in range(n):
                d[k - n + i] = (d[k - n + i] - t * self.low[i]) % q
        return tuple(d[:n])

    def inv(self, a):
        # solve (multiplication-by-a matrix) y = 1 over F_q
        q, n = self.q, self.n
        x = tuple(1 if i == 1 else 0 for i in range(n)) if n > 1 else None
        cols, cur = [], a
        for _ in range(n):
            cols.append(cur)
            cur = self.mul(cur, x) if n > 1 else cur
        M = [[cols[j][i] for j in range(n)] + [1 if i == 0 else 0] for i in range(n)]
        for c in range(n):
            piv = next(r for r in range(c, n) if M[r][c])
            M[c], M[piv] = M[piv], M[c]
            iv = pow(M[c][c], -1, q)
            M[c] = [v * iv % q for v in M[c]]
            for r in range(n):
                if r != c and M[r][c]:
                    f = M[r][c]
                    M[r] = [(v - f * w) % q for v, w in zip(M[r], M[c])]
        return tuple(M[i][n] for i in range(n))

    def embed(self, k):
        return (k % self.q,) + (0,) * (self.n - 1)


class _ExtCurve:
    """y^2 = x^3 + a x + b over an _ExtField, affine"""

    def __init__(self, K, a, b):
        self.K, self.a, self.b = K, a, b

    def add(self, P, Q):
        K = self.K
        if P is None:
            return Q
        if Q is None:
            return P
        (x1, y1), (x2, y2) = P, Q
        if x1 == x2:
            if K.add(y1, y2) == K.zero:
                return None
            x1s = K.mul(x1, x1)
            lam = K.mul(K.add(K.add(K.add(x1s, x1s), x1s), self.a), K.inv(K.add(y1, y1)))
        else:
            lam = K.mul(K.sub(y2, y1), K.inv(K.sub(x2, x1)))
        x3 = K.sub(K.sub(K.mul(lam, lam), x1), x2)
        return (x3, K.sub(K.mul(lam, K.sub(x1, x3)), y1))

    def on_curve(self, P):
        K = self.K
        x, y = P
        return K.mul(y, y) == K.add(K.mul(K.add(K.mul(x, x), self.a), x), self.b)


def _dec(K, bs, width=None):
    n = K.n
    width = width or len(bs) // (2 * n)
    xs = [int.from_bytes(bs[i * width:(i + 1) * width], "big") for i in range(2 * n)]
    return (tuple(xs[:n]), tuple(xs[n:]))


def _enc(P, width=20):
    # width = bytes per F_q coordinate (20 for f.param / d159.param, 19 for g149.param)
    return b"".join(c.to_bytes(width, "big") for c in P[0]) + b"".join(c.to_bytes(width, "big") for c in P[1])


def _walk(E, start, step, g):
    out, cur = [], start
    for _ in range(g):
        out.append(cur)
        cur = E.add(cur, step)
    return out


def type_fd_points(param: dict, seeds_g1, seeds_g2, g: int):
    """2 x g distinct points (wire bytes) for a type f or type d (k = 6) parameter set, grown
    from two seed points per group (wire bytes, e.g. reference fixtures)."""
    q = param["q"]
    Fq = _ExtField(q, [0])           # F_q itself as a degree-1 "extension": tuples of length 1
    if param["type"] == "f":
        E1 = _ExtCurve(Fq, Fq.zero, Fq.embed(param["b"]))
        K2 = _ExtField(q, [-param["beta"], 0])                       # s^2 = beta
        xi = ((-param["alpha0"]) % q, (-param["alpha1"]) % q)
        E2 = _ExtCurve(K2, K2.zero, K2.mul(xi, K2.embed(param["b"])))
    else:
        E1 = _ExtCurve(Fq, Fq.embed(param["a"]), Fq.embed(param["b"]))
        ncoef = 5 if param["type"] == "g" else 3
        K2 = _ExtField(q, [param["coeff%d" % i] for i in range(ncoef)])
        v = param["nqr"] % q
        E2 = _ExtCurve(K2, K2.embed(param["a"] * v * v), K2.embed(param["b"] * v * v * v))
    P0, G = (_dec(Fq, s) for s in seeds_g1[:2])
    Q0, H = (_dec(K2, s) for s in seeds_g2[:2])
    assert all(E1.on_curve(p) for p in (P0, G)) and all(E2.on_curve(p) for p in (Q0, H))
    wb = (q.bit_length() + 7) // 8
    return [_enc(p, wb) for p in _walk(E1, P0, G, g)], [_enc(p, wb) for p in _walk(E2, Q0, H, g)]
