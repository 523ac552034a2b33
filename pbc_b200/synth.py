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
    """2 x g distinct points of the order-r subgroup of y^2 = x^3 + x (Type A, G1 = G2)."""
    q, h = param["q"], param["h"]
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
    enc = lambda pt: pt[0].to_bytes(64, "big") + pt[1].to_bytes(64, "big")
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
