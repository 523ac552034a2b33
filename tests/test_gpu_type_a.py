"""GPU parity tests, Type A: CUDA path (through the C ABI) vs oracle, reference fixtures and
properties.  Integer work: every comparison is bit-exact."""
import os
import random

import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from pbc_b200.pairing import Pairing
    return Pairing(PARAMS["a"])


@pytest.fixture(scope="module")
def orc():
    return O.pairing_from_param(PARAMS["a"])


def _cat(xs):
    return b"".join(bytes.fromhex(x) for x in xs)


# ---- F_p layer: differential test against exact integer arithmetic (guru/fp_test.c:44-85) ----
@pytest.mark.parametrize("op", [0, 1, 2, 3, 4, 5, 6, 7])
def test_fp_ops_match_integers(dev, orc, op):
    q = orc.q
    rnd = random.Random(1234 + op)
    n = 1000
    a = [rnd.randrange(q) for _ in range(n)]
    b = [rnd.randrange(q) for _ in range(n)]
    # edge values: 0, 1, q-1, values next to 2^512 wraparound
    edge = [0, 1, 2, q - 1, q - 2, (q + 1) // 2, (1 << 511) % q, q >> 1]
    for i, e in enumerate(edge):
        a[i] = e
        b[i] = edge[(i * 3 + 1) % len(edge)]
    if op == 3:
        a = [x if x else 1 for x in a]
    A = b"".join(x.to_bytes(64, "big") for x in a)
    B = b"".join(x.to_bytes(64, "big") for x in b)
    got = dev.fp_op(op, A, B, n)
    f = {0: lambda x, y: x * y % q, 1: lambda x, y: (x + y) % q, 2: lambda x, y: (x - y) % q,
         3: lambda x, y: pow(x, -1, q), 4: lambda x, y: x * pow(2, -1, q) % q,
         5: lambda x, y: (-x) % q, 6: lambda x, y: x * x % q,
         7: lambda x, y: (x * y - y) % q}[op]
    want = b"".join(f(x, y).to_bytes(64, "big") for x, y in zip(a, b))
    assert got == want


def test_fp_accepts_unreduced_input(dev, orc):
    # fp_from_bytes reduces mod q (arith/montfp.c:498-517)
    q = orc.q
    vals = [q, q + 1, (1 << 512) - 1, (1 << 512) - q]
    A = b"".join(v.to_bytes(64, "big") for v in vals)
    B = b"".join((1).to_bytes(64, "big") for _ in vals)
    got = dev.fp_op(0, A, B, len(vals))
    assert got == b"".join((v % q).to_bytes(64, "big") for v in vals)


# ---- pairing: reference fixtures (compiled reference) and the reference's own KAT ----
def test_reference_fixtures(dev, golden):
    g = golden["a"]["pairing"]
    n = len(g["e"])
    assert dev.apply(_cat(g["P"]), _cat(g["Q"]), n) == _cat(g["e"])


def test_reference_kat_pairing_test_pbc(dev):
    # pbc/pairing_test.pbc:3-21
    from tests.test_oracle import KAT_G, KAT_H, KAT_E_GH, KAT_GA, KAT_HB, KAT_RES
    enc = lambda pt: pt[0].to_bytes(64, "big") + pt[1].to_bytes(64, "big")
    got = dev.apply(enc(KAT_G) + enc(KAT_GA), enc(KAT_H) + enc(KAT_HB), 2)
    assert got == enc(KAT_E_GH) + enc(KAT_RES)


def test_matches_oracle_on_seeded_points(dev, orc):
    rnd = random.Random(99)
    G = orc.E.from_bytes(bytes.fromhex(__import__("json").load(open(os.path.join(
        os.path.dirname(__file__), "golden", "a.json")))["pairing"]["P"][0]))
    n = 37   # ragged: not a multiple of the block size
    Ps = [orc.E.mul(rnd.randrange(1, orc.r), G) for _ in range(n)]
    Qs = [orc.E.mul(rnd.randrange(1, orc.r), G) for _ in range(n)]
    P = b"".join(orc.E.to_bytes(p) for p in Ps)
    Q = b"".join(orc.E.to_bytes(p) for p in Qs)
    assert dev.apply(P, Q, n) == O.pairing_batch(orc, P, Q, n)


def test_offcurve_and_identity_semantics(dev, golden):
    g = golden["a"]
    P0, Q0 = bytes.fromhex(g["pairing"]["P"][0]), bytes.fromhex(g["pairing"]["Q"][0])
    ident = bytes.fromhex(g["offcurve"]["identity"])
    badP, badQ = bytes.fromhex(g["offcurve"]["badP"]), bytes.fromhex(g["offcurve"]["badQ"])
    got = dev.apply(badP + P0 + P0, Q0 + badQ + Q0, 3)
    assert got[:128] == ident and got[128:256] == ident
    assert got[256:] == bytes.fromhex(g["pairing"]["e"][0])


def test_empty_and_single(dev, golden):
    g = golden["a"]["pairing"]
    assert dev.apply(b"", b"", 0) == b""
    assert dev.apply(bytes.fromhex(g["P"][3]), bytes.fromhex(g["Q"][3]), 1) == bytes.fromhex(g["e"][3])


def test_bilinearity_fixture(dev, golden):
    # e(P^a, Q) from the reference; and symmetric pairing e(P,Q) == e(Q,P) for type A
    g = golden["a"]
    n = len(g["pow"]["Pa"])
    got = dev.apply(_cat(g["pow"]["Pa"]), _cat(g["pairing"]["Q"][:n]), n)
    assert got == _cat(g["pow"]["e_Pa_Q"])
    sym = dev.apply(_cat(g["pairing"]["Q"][:n]), _cat(g["pairing"]["P"][:n]), n)
    assert sym == _cat(g["pairing"]["e"][:n])


def test_large_batch_properties(dev, golden):
    """BASELINE-size style check without the oracle: a multi-chunk batch built by tiling the
    fixtures must reproduce the fixture outputs at every position (determinism, chunking,
    multi-stream pipeline) -- 2^18 + 777 pairings crosses the chunk boundary raggedly."""
    g = golden["a"]["pairing"]
    m = len(g["e"])
    P, Q, E = _cat(g["P"]), _cat(g["Q"]), _cat(g["e"])
    n = (1 << 18) + 777
    reps = n // m + 1
    got = dev.apply((P * reps)[:n * 128], (Q * reps)[:n * 128], n)
    assert got == (E * reps)[:n * 128]


def test_device_pointer_entry(dev, golden):
    import torch
    g = golden["a"]["pairing"]
    n = len(g["e"])
    P = torch.frombuffer(bytearray(_cat(g["P"])), dtype=torch.uint8).cuda()
    Q = torch.frombuffer(bytearray(_cat(g["Q"])), dtype=torch.uint8).cuda()
    out = torch.empty(n * 128, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    dev.apply_device(out.data_ptr(), P.data_ptr(), Q.data_ptr(), n, st.cuda_stream)
    st.synchronize()
    assert bytes(out.cpu().numpy().tobytes()) == _cat(g["e"])


def test_device_pointer_calls_on_two_streams_share_the_workspace_in_order(dev, golden):
    """the _device entry points of one handle use one workspace per device; the library orders its users
    with an event, so back-to-back calls on DIFFERENT streams (pairings, then a product, then a G1 power
    and a pp handle) must each give the reference bytes -- unordered they would overwrite each other's
    Miller values"""
    import torch
    g = golden["a"]
    reps = 600                                              # long enough for the calls to overlap if they could
    n = len(g["pairing"]["e"])
    P = torch.frombuffer(bytearray(_cat(g["pairing"]["P"]) * reps), dtype=torch.uint8).cuda()
    Q = torch.frombuffer(bytearray(_cat(g["pairing"]["Q"]) * reps), dtype=torch.uint8).cuda()
    k, n_out = g["prod"]["k"], len(g["prod"]["e"])
    PP = torch.frombuffer(bytearray(_cat(g["prod"]["P"]) * reps), dtype=torch.uint8).cuda()
    QQ = torch.frombuffer(bytearray(_cat(g["prod"]["Q"]) * reps), dtype=torch.uint8).cuda()
    o1 = torch.empty(n * reps * 128, dtype=torch.uint8, device="cuda")
    o2 = torch.empty(n_out * reps * 128, dtype=torch.uint8, device="cuda")
    o3 = torch.empty(n * reps * 128, dtype=torch.uint8, device="cuda")
    o4 = torch.empty(n * reps * 128, dtype=torch.uint8, device="cuda")
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    h = dev.pp_init(bytes.fromhex(g["pairing"]["P"][0]))
    dev.apply_device(o1.data_ptr(), P.data_ptr(), Q.data_ptr(), n * reps, s1.cuda_stream)
    dev.prod_apply_device(o2.data_ptr(), PP.data_ptr(), QQ.data_ptr(), k, n_out * reps, s2.cuda_stream)
    dev.apply_device(o3.data_ptr(), P.data_ptr(), Q.data_ptr(), n * reps, s3.cuda_stream)
    h.apply_device(o4.data_ptr(), Q.data_ptr(), n * reps, s1.cuda_stream)
    torch.cuda.synchronize()
    assert bytes(o1.cpu().numpy().tobytes()) == _cat(g["pairing"]["e"]) * reps
    assert bytes(o2.cpu().numpy().tobytes()) == _cat(g["prod"]["e"]) * reps
    assert bytes(o3.cpu().numpy().tobytes()) == _cat(g["pairing"]["e"]) * reps
    want = dev.apply(bytes.fromhex(g["pairing"]["P"][0]) * n, _cat(g["pairing"]["Q"]), n)
    assert bytes(o4.cpu().numpy().tobytes()) == want * reps
    h.clear()


# ---- element_prod_pairing (ecc/a_param.c:1283-1383) ----
def test_prod_reference_fixtures(dev, golden):
    g = golden["a"]["prod"]
    k, n_out = g["k"], len(g["e"])
    assert dev.prod_apply(_cat(g["P"]), _cat(g["Q"]), k, n_out) == _cat(g["e"])


def test_prod_matches_oracle_and_product_of_singles(dev, orc, golden):
    g = golden["a"]["pairing"]
    k, n_out = 5, 4                       # 20 of the 24 fixture pairs, k not a power of two
    P, Q = _cat(g["P"][:k * n_out]), _cat(g["Q"][:k * n_out])
    got = dev.prod_apply(P, Q, k, n_out)
    for i in range(n_out):
        want = O.prod_pairing_bytes(orc, [bytes.fromhex(x) for x in g["P"][i * k:(i + 1) * k]],
                                    [bytes.fromhex(x) for x in g["Q"][i * k:(i + 1) * k]])
        assert got[i * 128:(i + 1) * 128] == want
        acc = orc.GT.one
        for j in range(i * k, (i + 1) * k):
            acc = orc.GT.mul(acc, orc.GT.from_bytes(bytes.fromhex(g["e"][j])))
        assert want == orc.GT.to_bytes(acc)


def test_prod_any_infinite_input_gives_identity(dev, golden):
    # include/pbc_pairing.h:161-168: one O anywhere makes the whole product 1
    g = golden["a"]
    k = g["prod"]["k"]
    P, Q = [bytes.fromhex(x) for x in g["prod"]["P"]], [bytes.fromhex(x) for x in g["prod"]["Q"]]
    P[1] = bytes.fromhex(g["offcurve"]["badP"])          # poisons output 0 only
    Q[2 * k + 3] = bytes.fromhex(g["offcurve"]["badQ"])  # poisons output 2 only
    got = dev.prod_apply(b"".join(P), b"".join(Q), k, 3)
    ident = bytes.fromhex(g["offcurve"]["identity"])
    assert got[:128] == ident and got[256:] == ident
    assert got[128:256] == bytes.fromhex(g["prod"]["e"][1])


@pytest.mark.parametrize("share", [2, 4, 3, 16])
def test_prod_with_a_shared_miller_accumulator(golden, orc, share):
    """a_pairings_affine shares one accumulator between the pairs of a product (ecc/a_param.c:1296-1306);
    k_a_miller9_shared does so for `share` pairs per thread ("b200_prod_share" sets it; the largest divisor of
    k not above it is used): the reference fixtures (k = 4), an O somewhere, and k = 12 / k = 16 products of the
    fixture pairs against the oracle"""
    from pbc_b200.pairing import Pairing
    dev = Pairing(PARAMS["a"] + "\nb200_prod_share %d\n" % share)
    g = golden["a"]
    k, n_out = g["prod"]["k"], len(g["prod"]["e"])
    assert dev.prod_apply(_cat(g["prod"]["P"]), _cat(g["prod"]["Q"]), k, n_out) == _cat(g["prod"]["e"])
    P, Q = [bytes.fromhex(x) for x in g["prod"]["P"]], [bytes.fromhex(x) for x in g["prod"]["Q"]]
    P[k + 2] = bytes.fromhex(g["offcurve"]["badP"])          # poisons output 1 only
    got = dev.prod_apply(b"".join(P), b"".join(Q), k, n_out)
    assert got[128:256] == bytes.fromhex(g["offcurve"]["identity"]) and got[:128] == bytes.fromhex(g["prod"]["e"][0])
    pairs = g["pairing"]
    for kk in (12, 16):
        from pbc_b200 import _lib
        n2 = 3 if _lib.IS_SIMULATOR else 130                 # on the GPU: more than one thread block at share = 1
        idx = [(i * 7 + j * 5) % len(pairs["e"]) for i in range(n2) for j in range(kk)]
        Pb = b"".join(bytes.fromhex(pairs["P"][t]) for t in idx)
        Qb = b"".join(bytes.fromhex(pairs["Q"][t]) for t in idx)
        got = dev.prod_apply(Pb, Qb, kk, n2)
        for i in sorted({0, 1, min(64, n2 - 1), n2 - 1}):
            acc = orc.GT.one
            for t in idx[i * kk:(i + 1) * kk]:
                acc = orc.GT.mul(acc, orc.GT.from_bytes(bytes.fromhex(pairs["e"][t])))
            assert got[i * 128:(i + 1) * 128] == orc.GT.to_bytes(acc), (kk, i)


def test_prod_k1_equals_single_and_empty(dev, golden):
    g = golden["a"]["pairing"]
    assert dev.prod_apply(_cat(g["P"]), _cat(g["Q"]), 1, len(g["e"])) == _cat(g["e"])
    assert dev.prod_apply(b"", b"", 4, 0) == b""


def test_prod_large_batch_tiles(dev, golden):
    """2^12+5 outputs of k = 16 (config 5 shape) built by tiling the fixture pairs: crosses the
    host pipeline's chunk boundary; every output must equal the oracle product of its 16 pairs."""
    g = golden["a"]["pairing"]
    m = len(g["e"])                       # 24 pairs: period lcm(24,16)=48 pairs = 3 outputs
    k, n_out = 16, (1 << 14) + 5
    reps = k * n_out // m + 1
    P, Q = (_cat(g["P"]) * reps)[:k * n_out * 128], (_cat(g["Q"]) * reps)[:k * n_out * 128]
    got = dev.prod_apply(P, Q, k, n_out)
    first = dev.prod_apply(P[:48 * 128], Q[:48 * 128], k, 3)
    assert got == (first * (n_out // 3 + 1))[:n_out * 128]


# ---- pairing_pp_init / pairing_pp_apply (ecc/a_param.c:149-220, 317-360) ----
def test_pp_reference_fixtures(dev, golden):
    g = golden["a"]
    n = len(g["pp"]["e"])
    got = dev.pp_apply(bytes.fromhex(g["pp"]["P"]), _cat(g["pairing"]["Q"][:n]), n)
    assert got == _cat(g["pp"]["e"])


def test_pp_equals_plain_pairing(dev, golden):
    g = golden["a"]["pairing"]
    n = len(g["e"])
    P3 = bytes.fromhex(g["P"][3])
    assert dev.pp_apply(P3, _cat(g["Q"]), n) == dev.apply(P3 * n, _cat(g["Q"]), n)


def test_pp_offcurve(dev, golden):
    g = golden["a"]
    ident = bytes.fromhex(g["offcurve"]["identity"])
    Q = _cat(g["pairing"]["Q"][:2]) + bytes.fromhex(g["offcurve"]["badQ"])
    got = dev.pp_apply(bytes.fromhex(g["pp"]["P"]), Q, 3)
    assert got[256:] == ident and got[:256] == _cat(g["pp"]["e"][:2])
    assert dev.pp_apply(bytes.fromhex(g["offcurve"]["badP"]), Q, 3) == ident * 3
