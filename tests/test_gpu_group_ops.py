"""GPU parity tests of the operations either side of the pairing: batched
element_pow_zn on G1 and on GT for types A, F, D, against the reference fixtures (P^a, e(P,Q)^a
from the compiled reference) and the oracle's exact arithmetic.  Bit-exact."""
import random

import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

pytestmark = pytest.mark.gpu

VARIANTS = {"a": ("a", ""), "f": ("f", ""), "f_refbasis": ("f", "b200_reference_basis 1\n"), "d159": ("d159", ""),
            "d159_refbasis": ("d159", "b200_reference_basis 1\n"),
            "g149": ("g149", "")}


@pytest.fixture(scope="module", params=sorted(VARIANTS))
def env(request, golden):
    from pbc_b200.pairing import Pairing
    name, extra = VARIANTS[request.param]
    return {"dev": Pairing(PARAMS[name] + extra), "orc": O.pairing_from_param(PARAMS[name]), "g": golden[name]}


def _cat(xs):
    return b"".join(bytes.fromhex(x) for x in xs)


def test_zr_length(env):
    assert env["dev"].zr_len == env["g"]["lengths"]["zr"] == (env["orc"].r.bit_length() + 7) // 8


def test_g1_pow_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    assert d.g1_pow_zn(_cat(g["pairing"]["P"][:n]), _cat(g["pow"]["a"]), n) == _cat(g["pow"]["Pa"])


def test_g1_pow_matches_oracle_edge_scalars(env):
    g, d, orc = env["g"], env["dev"], env["orc"]
    rnd = random.Random(21)
    r = orc.r
    ks = [0, 1, 2, 3, r - 1, r - 2, r, r + 5, (1 << 160) - 1, 1 << 159, 0x80000000, 0xFFFFFFFF] + \
         [rnd.randrange(r) for _ in range(28)]
    n = len(ks)
    pts = [bytes.fromhex(g["pairing"]["P"][i % len(g["pairing"]["P"])]) for i in range(n)]
    got = d.g1_pow_zn(b"".join(pts), b"".join((k % (1 << (8 * d.zr_len))).to_bytes(d.zr_len, "big") for k in ks), n)
    L = d.g1_len
    for i, (k, pb) in enumerate(zip(ks, pts)):
        R = orc.G1.mul((k % (1 << (8 * d.zr_len))) % r, orc.G1.from_bytes(pb))
        want = bytes(L) if R is None else orc.G1.to_bytes(R)
        assert got[i * L:(i + 1) * L] == want, "scalar %d" % i


def test_g1_pow_offcurve_is_infinity_and_empty(env):
    g, d = env["g"], env["dev"]
    bad = bytes.fromhex(g["offcurve"]["badP"])
    assert d.g1_pow_zn(bad, (7).to_bytes(d.zr_len, "big"), 1) == bytes(d.g1_len)
    assert d.g1_pow_zn(b"", b"", 0) == b""


def test_gt_pow_reference_fixtures(env):
    """e(P, Q)^a == e(P^a, Q), both sides from the compiled reference"""
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    assert d.gt_pow_zn(_cat(g["pairing"]["e"][:n]), _cat(g["pow"]["a"]), n) == _cat(g["pow"]["e_Pa_Q"])


def test_gt_pow_matches_oracle_edge_scalars(env):
    g, d, orc = env["g"], env["dev"], env["orc"]
    rnd = random.Random(22)
    r = orc.r
    ks = [0, 1, 2, r - 1, r, r + 3, 1 << 159, 0xFFFFFFFF00000000] + [rnd.randrange(r) for _ in range(8)]
    n = len(ks)
    es = [bytes.fromhex(g["pairing"]["e"][i % len(g["pairing"]["e"])]) for i in range(n)]
    got = d.gt_pow_zn(b"".join(es), b"".join((k % (1 << (8 * d.zr_len))).to_bytes(d.zr_len, "big") for k in ks), n)
    L = d.gt_len
    for i, (k, eb) in enumerate(zip(ks, es)):
        kk = (k % (1 << (8 * d.zr_len))) % r
        want = orc.GT.to_bytes(orc.GT.pow(orc.GT.from_bytes(eb), kk)) if kk else orc.GT.to_bytes(orc.GT.one)
        assert got[i * L:(i + 1) * L] == want, "scalar %d" % i


def test_pow_then_pair_is_bilinear_on_device(env):
    """e(aP, Q) computed entirely on the GPU equals e(P, Q)^a computed entirely on the GPU."""
    g, d = env["g"], env["dev"]
    rnd = random.Random(23)
    n = 8
    ks = b"".join(rnd.randrange(1, env["orc"].r).to_bytes(d.zr_len, "big") for _ in range(n))
    P, Q = _cat(g["pairing"]["P"][:n]), _cat(g["pairing"]["Q"][:n])
    lhs = d.apply(d.g1_pow_zn(P, ks, n), Q, n)
    rhs = d.gt_pow_zn(d.apply(P, Q, n), ks, n)
    assert lhs == rhs


# ---- G2: the twist over F_q^2 (type f) / F_q^3 (type d); type a: G2 = G1 ----
def test_g2_pow_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    assert d.g2_pow_zn(_cat(g["pairing"]["Q"][:n]), _cat(g["pow"]["a"]), n) == _cat(g["pow"]["Qa"])


def test_g2_pow_matches_oracle_edge_scalars(env):
    g, d, orc = env["g"], env["dev"], env["orc"]
    rnd = random.Random(31)
    r = orc.r
    ks = [0, 1, 2, 3, r - 1, r, r + 5, 1 << 159, 0xFFFFFFFF] + [rnd.randrange(r) for _ in range(7)]
    n = len(ks)
    pts = [bytes.fromhex(g["pairing"]["Q"][i % len(g["pairing"]["Q"])]) for i in range(n)]
    got = d.g2_pow_zn(b"".join(pts), b"".join((k % (1 << (8 * d.zr_len))).to_bytes(d.zr_len, "big") for k in ks), n)
    L = d.g2_len
    for i, (k, pb) in enumerate(zip(ks, pts)):
        R = orc.G2.mul((k % (1 << (8 * d.zr_len))) % r, orc.G2.from_bytes(pb))
        want = bytes(L) if R is None else orc.G2.to_bytes(R)
        assert got[i * L:(i + 1) * L] == want, "scalar %d" % i
    bad = bytes.fromhex(g["offcurve"]["badQ"])
    assert d.g2_pow_zn(bad, (5).to_bytes(d.zr_len, "big"), 1) == bytes(L)


def test_both_arguments_powered_on_device(env):
    """e(aP, aQ) with both multiples taken on the GPU == the reference's e(P^a, Q^a)"""
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    a = _cat(g["pow"]["a"])
    aP = d.g1_pow_zn(_cat(g["pairing"]["P"][:n]), a, n)
    aQ = d.g2_pow_zn(_cat(g["pairing"]["Q"][:n]), a, n)
    assert d.apply(aP, aQ, n) == _cat(g["pow"]["e_Pa_Qa"])


# ---- element_from_hash on G1 (ecc/curve.c:455-482) ----
@pytest.mark.parametrize("length", [3, 20, 32, 70])
def test_g1_from_hash_reference_fixtures(env, length):
    blk = env["g"]["hash"][str(length)]
    got = env["dev"].g1_from_hash(_cat(blk["data"]), length, len(blk["data"]))
    assert got == _cat(blk["G1"])


def test_g1_from_hash_matches_oracle_on_many_inputs(env):
    import hashlib
    d, orc = env["dev"], env["orc"]
    n = 40
    hs = [hashlib.sha256(b"msg-%d" % i).digest() for i in range(n)]
    got = d.g1_from_hash(b"".join(hs), 32, n)
    L = d.g1_len
    for i, h in enumerate(hs):
        assert got[i * L:(i + 1) * L] == orc.G1.to_bytes(O.g1_from_hash(orc, h)), i


def test_bls_style_verification_entirely_on_device(env):
    """sigma = sk * H(m), pk = sk * g2: e(sigma, g2) == e(H(m), pk) -- hash-to-curve, both scalar
    multiplications and both pairings on the GPU; a forged signature must fail."""
    import hashlib
    g, d, orc = env["g"], env["dev"], env["orc"]
    n = 6
    rnd = random.Random(41)
    sk = rnd.randrange(1, orc.r).to_bytes(d.zr_len, "big")
    g2 = bytes.fromhex(g["pairing"]["Q"][0])
    pk = d.g2_pow_zn(g2, sk, 1)
    msgs = b"".join(hashlib.sha256(b"bls-%d" % i).digest() for i in range(n))
    H = d.g1_from_hash(msgs, 32, n)
    sig = d.g1_pow_zn(H, sk * n, n)
    lhs = d.apply(sig, g2 * n, n)
    rhs = d.apply(H, pk * n, n)
    assert lhs == rhs
    forged = d.g1_pow_zn(H, (int.from_bytes(sk, "big") ^ 1).to_bytes(d.zr_len, "big") * n, n)
    assert d.apply(forged, g2 * n, n) != rhs


# ---- element_to/from_bytes_compressed on G1 (ecc/curve.c:762-813) ----
def test_g1_compressed_roundtrip_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["compressed"]["G1"])
    P = _cat(g["pairing"]["P"][:n])
    comp = _cat(g["compressed"]["G1"])
    assert d.g1_compress(P, n) == comp                      # host-side compression == the reference's
    assert d.g1_decompress(comp, n) == P
    clen = g["compressed"]["len"]
    flipped = b"".join(comp[i * clen:(i + 1) * clen - 1] + bytes([comp[(i + 1) * clen - 1] ^ 1]) for i in range(n))
    assert d.g1_decompress(flipped, n) == _cat(g["compressed"]["G1_flipped_sign"])


def test_g1_decompress_x_without_point_gives_zero_bytes(env):
    d, orc = env["dev"], env["orc"]
    q, E = orc.q, orc.G1
    x = 2
    while pow(((x * x + E.a) * x + E.b) % q, (q - 1) // 2, q) != q - 1:
        x += 1
    L = d.g1_len
    assert d.g1_decompress(x.to_bytes(L // 2, "big") + b"\x01", 1) == bytes(L)
