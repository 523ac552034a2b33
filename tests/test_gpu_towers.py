"""GPU differential tests of the extension towers of Types F and D (F_q^2/F_q^12, F_q^3/F_q^6):
CUDA tower routines through pbc_b200_tower_op vs the oracle's exact arithmetic on the same random
GT-sized operands.  Bit-exact."""
import random

import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

pytestmark = pytest.mark.gpu


def _rand_gt(orc, rnd, n):
    q = orc.q
    k = orc.gt_len // 20
    return [[rnd.randrange(q) for _ in range(k)] for _ in range(n)]


def _bytes(elems):
    return b"".join(b"".join(c.to_bytes(20, "big") for c in e) for e in elems)


@pytest.fixture(scope="module", params=["internal_basis", "reference_basis"])
def f_env(request):
    from pbc_b200.pairing import Pairing
    extra = "" if request.param == "internal_basis" else "b200_reference_basis 1\n"
    return Pairing(PARAMS["f"] + extra), O.pairing_from_param(PARAMS["f"])


@pytest.fixture(scope="module", params=["internal_basis", "reference_basis"])
def d_env(request):
    from pbc_b200.pairing import Pairing
    extra = "" if request.param == "internal_basis" else "b200_reference_basis 1\n"
    return Pairing(PARAMS["d159"] + extra), O.pairing_from_param(PARAMS["d159"])


def _f12(e):
    return tuple((e[2 * i], e[2 * i + 1]) for i in range(6))


@pytest.mark.parametrize("op", [0, 1, 2, 4, 3])
def test_f12_ops(f_env, op):
    dev, orc = f_env
    rnd = random.Random(100 + op)
    n = 8 if op == 3 else 64
    A, B = _rand_gt(orc, rnd, n), _rand_gt(orc, rnd, n)
    A[0] = [1] + [0] * 11
    if op != 2:
        A[1] = [0] * 12
    got = dev.tower_op(op, _bytes(A), _bytes(B), n)
    F12, F2 = orc.Fq12, orc.Fq2
    want = []
    for a, b in zip(A, B):
        x, y = _f12(a), _f12(b)
        if op == 0:
            r = F12.mul(x, y)
        elif op == 1:
            r = F12.sqr(x)
        elif op == 2:
            r = F12.inv(x)
        elif op == 3:
            r = orc.final_exp(x) if any(a) else None
        else:
            line = [F2.zero] * 6
            line[0] = (y[0][0], 0)
            line[3], line[4] = y[3], y[4]
            r = F12.mul(x, tuple(line))
        want.append(r)
    L = orc.gt_len
    for i, r in enumerate(want):
        if r is None:
            continue           # final exponentiation of 0: the reference would divide by zero
        assert got[i * L:(i + 1) * L] == F12.to_bytes(r), "op %d element %d" % (op, i)


def test_f12_cyclotomic_square(f_env):
    """Granger-Scott squaring (tower op 5) on elements of the cyclotomic subgroup == generic square"""
    dev, orc = f_env
    rnd = random.Random(77)
    F12, F2, q = orc.Fq12, orc.Fq2, orc.q
    elems = []
    for _ in range(16):
        f = tuple((rnd.randrange(q), rnd.randrange(q)) for _ in range(6))
        conj = tuple(c if i % 2 == 0 else F2.neg(c) for i, c in enumerate(f))
        g = F12.mul(conj, F12.inv(f))
        elems.append(F12.mul(F12.pow(g, q * q), g))
    A = b"".join(F12.to_bytes(g) for g in elems)
    got = dev.tower_op(5, A, A, len(elems))
    L = orc.gt_len
    for i, g in enumerate(elems):
        assert got[i * L:(i + 1) * L] == F12.to_bytes(F12.sqr(g)), i


def _f6d(e):
    return (tuple(e[:3]), tuple(e[3:]))


@pytest.mark.parametrize("op", [5, 6, 0, 1, 2, 3])
def test_f6d_ops(d_env, op):
    dev, orc = d_env
    rnd = random.Random(200 + op)
    n = 16 if op == 3 else 64
    A, B = _rand_gt(orc, rnd, n), _rand_gt(orc, rnd, n)
    A[0] = [1] + [0] * 5
    got = dev.tower_op(op, _bytes(A), _bytes(B), n)
    F3, F6 = orc.Fq3, orc.Fq6
    L = orc.gt_len
    for i, (a, b) in enumerate(zip(A, B)):
        x, y = _f6d(a), _f6d(b)
        if op == 0:
            r = F6.mul(x, y)
        elif op == 1:
            r = F6.sqr(x)
        elif op == 2:
            r = F6.inv(x)
        elif op == 3:
            r = orc.tatepower(x)
        elif op == 5:
            r = (F3.mul(x[0], y[0]), F3.zero)
        else:
            r = (F3.inv(x[0]), F3.zero)
        assert got[i * L:(i + 1) * L] == F6.to_bytes(r), "op %d element %d" % (op, i)


@pytest.fixture(scope="module")
def g_env():
    from pbc_b200.pairing import Pairing
    return Pairing(PARAMS["g149"]), O.pairing_from_param(PARAMS["g149"])


@pytest.mark.parametrize("op", [5, 6, 0, 1, 2, 3])
def test_f10_ops(g_env, op):
    """F_q^5 / F_q^10 of type g (19-byte coordinates) vs the oracle"""
    dev, orc = g_env
    rnd = random.Random(300 + op)
    q = orc.q
    n = 6 if op == 3 else 32
    A = [[rnd.randrange(q) for _ in range(10)] for _ in range(n)]
    B = [[rnd.randrange(q) for _ in range(10)] for _ in range(n)]
    A[0] = [1] + [0] * 9
    enc = lambda es: b"".join(b"".join(c.to_bytes(19, "big") for c in e) for e in es)
    got = dev.tower_op(op, enc(A), enc(B), n)
    F5, F10 = orc.Fq5, orc.Fq10
    L = orc.gt_len
    for i, (a, b) in enumerate(zip(A, B)):
        x, y = (tuple(a[:5]), tuple(a[5:])), (tuple(b[:5]), tuple(b[5:]))
        if op == 0:
            r = F10.mul(x, y)
        elif op == 1:
            r = F10.sqr(x)
        elif op == 2:
            r = F10.inv(x)
        elif op == 3:
            r = orc.tatepower(x)
        elif op == 5:
            r = (F5.mul(x[0], y[0]), F5.zero)
        else:
            r = (F5.inv(x[0]), F5.zero)
        assert got[i * L:(i + 1) * L] == F10.to_bytes(r), "op %d element %d" % (op, i)
