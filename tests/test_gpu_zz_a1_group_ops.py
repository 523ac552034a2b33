"""The operations either side of the Type A1 pairing on the GPU (group_a1.cuh): element_pow_zn on
G1 = G2 and on GT, element_from_hash and compressed points on G1, for the 1033-bit a1.param --
byte-for-byte against fixtures of the compiled reference (tests/golden/a1.json) and the oracle.
(Named to run after the other GPU tests: these kernels were first validated on the CPU simulator of
the library, tests/test_kernels_on_cpu_sim.py.)"""
import json
import os

import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cat(xs):
    return b"".join(bytes.fromhex(x) for x in xs)


@pytest.fixture(scope="module")
def env():
    from pbc_b200.pairing import Pairing
    with open(os.path.join(ROOT, "tests", "golden", "a1.json")) as f:
        g = json.load(f)
    dev = Pairing(PARAMS["a1"])
    yield {"dev": dev, "orc": O.pairing_from_param(PARAMS["a1"]), "g": g}
    dev.clear()


def test_g1_and_g2_powers_match_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    a = _cat(g["pow"]["a"])
    assert d.g1_pow_zn(_cat(g["pairing"]["P"][:n]), a, n) == _cat(g["pow"]["Pa"])
    assert d.g2_pow_zn(_cat(g["pairing"]["Q"][:n]), a, n) == _cat(g["pow"]["Qa"])


def test_gt_power_matches_pairing_of_the_multiple(env):
    """e(P, Q)^a == e(aP, Q) (fixture e_Pa_Q), the power taken on the GPU"""
    g, d = env["g"], env["dev"]
    n = len(g["pow"]["a"])
    assert d.gt_pow_zn(_cat(g["pairing"]["e"][:n]), _cat(g["pow"]["a"]), n) == _cat(g["pow"]["e_Pa_Q"])


def test_edge_scalars(env):
    """0 -> O (zero bytes) / 1 in GT; 1; n - 1; n (= 0); values above n are reduced"""
    d, orc, g = env["dev"], env["orc"], env["g"]
    P0 = bytes.fromhex(g["pairing"]["P"][0])
    E0 = bytes.fromhex(g["pairing"]["e"][0])
    Pp, e0 = orc.G1.from_bytes(P0), orc.GT.from_bytes(E0)
    ks = [0, 1, 2, orc.r - 1, orc.r, orc.r + 5, (1 << (8 * d.zr_len)) - 1]
    K = b"".join(k.to_bytes(d.zr_len, "big") for k in ks)
    got = d.g1_pow_zn(P0 * len(ks), K, len(ks))
    L = d.g1_len
    for i, k in enumerate(ks):
        w = orc.E.mul(k % orc.r, Pp)
        assert got[i * L:(i + 1) * L] == (bytes(L) if w is None else orc.G1.to_bytes(w)), k
    got = d.gt_pow_zn(E0 * len(ks), K, len(ks))
    for i, k in enumerate(ks):
        assert got[i * L:(i + 1) * L] == orc.GT.to_bytes(orc.GT.pow(e0, k % orc.r)), k


def test_power_then_pair_is_bilinear_on_device(env):
    g, d = env["g"], env["dev"]
    n = 2
    a = _cat(g["pow"]["a"][:n])
    aP = d.g1_pow_zn(_cat(g["pairing"]["P"][:n]), a, n)
    aQ = d.g2_pow_zn(_cat(g["pairing"]["Q"][:n]), a, n)
    assert d.apply(aP, aQ, n) == _cat(g["pow"]["e_Pa_Qa"][:n])


@pytest.mark.parametrize("length", [3, 20, 32, 70])
def test_g1_from_hash_reference_fixtures(env, length):
    blk = env["g"]["hash"][str(length)]
    got = env["dev"].g1_from_hash(_cat(blk["data"]), length, len(blk["data"]))
    assert got == _cat(blk["G1"])


def test_g1_compressed_roundtrip_reference_fixtures(env):
    g, d = env["g"], env["dev"]
    n = len(g["compressed"]["G1"])
    P = _cat(g["pairing"]["P"][:n])
    comp = _cat(g["compressed"]["G1"])
    clen = g["compressed"]["len"]
    assert d.g1_compress(P, n) == comp
    assert d.g1_decompress(comp, n) == P
    flipped = b"".join(comp[i * clen:(i + 1) * clen - 1] + bytes([comp[(i + 1) * clen - 1] ^ 1]) for i in range(n))
    assert d.g1_decompress(flipped, n) == _cat(g["compressed"]["G1_flipped_sign"])


def test_decompress_x_without_point_gives_zero_bytes(env):
    d, orc = env["dev"], env["orc"]
    p = orc.q
    x = 2
    while pow((x * x * x + x) % p, (p - 1) // 2, p) != p - 1:
        x += 1
    wb = d.g1_len // 2
    assert d.g1_decompress(x.to_bytes(wb, "big") + b"\x01", 1) == bytes(d.g1_len)
