"""Type A1 (ecc/a_param.c:1564-2273) on the GPU: the 34-limb field, pairings, products of pairings
and fixed-argument pairings, byte-for-byte against fixtures the compiled reference produced
(tests/golden/a1.json: param/a1.param, 1033-bit p; tests/golden/a1_small.json: a 167-bit p made the
way pbc_param_init_a1_gen makes one) and against the oracle."""
import json
import os
import random

import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cat(xs):
    return b"".join(bytes.fromhex(x) for x in xs)


@pytest.fixture(scope="module", params=["a1", "a1_small"])
def env(request):
    from pbc_b200.pairing import Pairing
    with open(os.path.join(ROOT, "tests", "golden", request.param + ".json")) as f:
        g = json.load(f)
    text = PARAMS["a1"] if request.param == "a1" else g["param_text"]
    dev = Pairing(text)
    yield {"dev": dev, "orc": O.pairing_from_param(text), "g": g, "name": request.param}
    dev.clear()


def test_lengths(env):
    d, L = env["dev"], env["g"]["lengths"]
    assert (d.type, d.g1_len, d.g2_len, d.gt_len, d.zr_len) == ("a1", L["g1"], L["g2"], L["gt"], L["zr"])


# ---- F_p arithmetic in 34 limbs (analogue of guru/fp_test.c) ----
@pytest.mark.parametrize("op", [0, 1, 2, 4, 5, 6, 7, 3])
def test_field_ops_match_integers(env, op):
    d, p = env["dev"], env["orc"].q
    wb = d.g1_len // 2
    rnd = random.Random(100 + op)
    n = 70 if op != 3 else 20
    xs = [rnd.randrange(p) for _ in range(n)]
    ys = [rnd.randrange(p) for _ in range(n)]
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, 1 << (p.bit_length() - 1)]
    for i, e in enumerate(edge):
        xs[i] = e
        ys[i] = edge[(i * 3 + 1) % len(edge)]
    if op == 3:
        xs = [x or 1 for x in xs]
    A = b"".join(x.to_bytes(wb, "big") for x in xs)
    B = b"".join(y.to_bytes(wb, "big") for y in ys)
    got = d.fp_op(op, A, B, n)
    want = {0: lambda x, y: x * y % p, 1: lambda x, y: (x + y) % p, 2: lambda x, y: (x - y) % p,
            3: lambda x, y: pow(x, p - 2, p), 4: lambda x, y: x * pow(2, p - 2, p) % p,
            5: lambda x, y: -x % p, 6: lambda x, y: x * x % p, 7: lambda x, y: (x * y - y) % p}[op]
    for i in range(n):
        assert int.from_bytes(got[i * wb:(i + 1) * wb], "big") == want(xs[i], ys[i]), (op, i)


def test_non_canonical_wire_values_are_reduced(env):
    """fp_set_mpz reduces modulo p (arith/montfp.c:100-110): x + p on the wire behaves as x"""
    d, p = env["dev"], env["orc"].q
    wb = d.g1_len // 2
    xs = [v for v in (5, p - 3, 12345678901234567890) if v + p < (1 << (8 * wb))]
    A = b"".join((x + p).to_bytes(wb, "big") for x in xs)
    B = b"".join((7).to_bytes(wb, "big") for _ in xs)
    got = d.fp_op(0, A, B, len(xs))
    for i, x in enumerate(xs):
        assert int.from_bytes(got[i * wb:(i + 1) * wb], "big") == x * 7 % p


# ---- element_pairing ----
def test_pairings_match_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["pairing"]["e"])
    assert d.apply(_cat(g["pairing"]["P"]), _cat(g["pairing"]["Q"]), n) == _cat(g["pairing"]["e"])


def test_pairings_across_blocks_match_oracle(env):
    """every (P_i, Q_j) combination, repeated so the batch spans several thread blocks"""
    g, d, orc = env["g"], env["dev"], env["orc"]
    m = 4 if env["name"] == "a1" else 6
    Ps = [bytes.fromhex(x) for x in g["pairing"]["P"][:m]]
    Qs = [bytes.fromhex(x) for x in g["pairing"]["Q"][:m]]
    pairs = [(P, Q) for P in Ps for Q in Qs]
    want = [O.pairing_bytes(orc, P, Q) for P, Q in pairs]
    reps = 300 // len(pairs) + 1
    P = b"".join(p for p, _ in pairs) * reps
    Q = b"".join(q for _, q in pairs) * reps
    got = d.apply(P, Q, len(pairs) * reps)
    assert got == b"".join(want) * reps


def test_bilinearity_in_the_second_argument(env):
    """e(P, Q1) e(P, Q2) == e(P, Q1 + Q2), the sum taken by the oracle's curve arithmetic"""
    g, d, orc = env["g"], env["dev"], env["orc"]
    P = bytes.fromhex(g["pairing"]["P"][0])
    Q1 = orc.G2.from_bytes(bytes.fromhex(g["pairing"]["Q"][1]))
    Q2 = orc.G2.from_bytes(bytes.fromhex(g["pairing"]["Q"][2]))
    S = orc.G2.to_bytes(orc.E.add(Q1, Q2))
    out = d.apply(P * 3, orc.G2.to_bytes(Q1) + orc.G2.to_bytes(Q2) + S, 3)
    L = d.gt_len
    e1, e2, es = (orc.GT.from_bytes(out[i * L:(i + 1) * L]) for i in range(3))
    assert orc.GT.mul(e1, e2) == es


def test_offcurve_inputs_give_identity(env):
    """bytes that are not a point decode to O (ecc/curve.c:611-623) and e(O, .) = e(., O) = 1"""
    g, d = env["g"], env["dev"]
    ident = bytes.fromhex(g["offcurve"]["identity"])
    P0, Q0 = bytes.fromhex(g["pairing"]["P"][0]), bytes.fromhex(g["pairing"]["Q"][0])
    bad = bytes.fromhex(g["offcurve"]["badP"])      # G1 = G2: the same bytes are a bad Q as well
    out = d.apply(bad + P0 + P0, Q0 + bad + Q0, 3)
    assert out[:2 * d.gt_len] == ident * 2
    assert out[2 * d.gt_len:] == bytes.fromhex(g["pairing"]["e"][0])


# ---- element_prod_pairing ----
def test_products_match_reference_fixtures(env):
    g, d = env["g"]["prod"], env["dev"]
    k, n_out = g["k"], len(g["e"])
    assert d.prod_apply(_cat(g["P"][:k * n_out]), _cat(g["Q"][:k * n_out]), k, n_out) == _cat(g["e"])


def test_product_with_a_bad_input_is_identity(env):
    g, d = env["g"], env["dev"]
    k = g["prod"]["k"]
    P = bytearray(_cat(g["prod"]["P"][:k]))
    P[d.g1_len] ^= 1                               # second point leaves the curve
    out = d.prod_apply(bytes(P), _cat(g["prod"]["Q"][:k]), k, 1)
    assert out == bytes.fromhex(g["offcurve"]["identity"])


def test_product_of_one_equals_pairing(env):
    g, d = env["g"], env["dev"]
    n = 3
    P, Q = _cat(g["pairing"]["P"][:n]), _cat(g["pairing"]["Q"][:n])
    assert d.prod_apply(P, Q, 1, n) == _cat(g["pairing"]["e"][:n])


# ---- pairing_pp_init / pairing_pp_apply ----
def test_fixed_argument_pairings_match_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["pp"]["e"])
    assert d.pp_apply(bytes.fromhex(g["pp"]["P"]), _cat(g["pairing"]["Q"][:n]), n) == _cat(g["pp"]["e"])


def test_fixed_argument_with_bad_point_is_identity(env):
    g, d = env["g"], env["dev"]
    out = d.pp_apply(bytes.fromhex(g["offcurve"]["badP"]), _cat(g["pairing"]["Q"][:2]), 2)
    assert out == bytes.fromhex(g["offcurve"]["identity"]) * 2
