"""CPU-side checks of the drop-in boundary: libpbc_b200.so loads and exports every symbol that
include/pbc_b200.h declares; parameter parsing mirrors the reference's failure cases.  No GPU
compute is attempted here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from pbc_b200 import pairing
    return pairing


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "pbc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(pbc_b200_[A-Za-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, "pbc_b200", "libpbc_b200.so"))
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_init_accepts_standard_params_and_reports_lengths(built):
    from pbc_b200.params import PARAMS
    p = built.Pairing(PARAMS["a"])
    assert (p.type, p.g1_len, p.g2_len, p.gt_len) == ("a", 128, 128, 128)
    p.clear()


def test_init_failure_cases_match_reference(built):
    # ecc/param.c:171-198 (unknown type), :134-140 (missing key): init returns 1
    from pbc_b200.params import PARAMS
    with pytest.raises(built.PairingError):
        built.Pairing("q 17\nr 3\n")
    with pytest.raises(built.PairingError):
        built.Pairing("type zz\nq 17\n")
    broken = "\n".join(l for l in PARAMS["a"].splitlines() if not l.startswith("h "))
    with pytest.raises(built.PairingError, match="missing param"):
        built.Pairing(broken)
    with pytest.raises(built.PairingError):
        built.Pairing(PARAMS["a"].replace("exp2 159", "exp2 15x9").replace("r 7307", "r 8307"))


def test_parameter_ranges_this_build_cannot_represent_are_rejected(built):
    """a group order wider than the 160-bit scalar registers, or a modulus whose wire width
    (ceil(bits / 8), arith/montfp.c:577) is not the one the kernels are compiled for, must fail at
    init instead of producing a non-reference byte format"""
    from pbc_b200.params import PARAMS
    from pbc_b200 import synth
    A = synth.parse_param(PARAMS["a"])
    # r' = r * 2^40 (201 bits) with h' = h / 2^40 keeps r h = q + 1 when 2^40 | h; a.param's h is
    # divisible by 4 only, so craft the text directly: the check on r comes before r h == q + 1
    big_r = PARAMS["a"].replace("r %d" % A["r"], "r %d" % (A["r"] << 41))
    with pytest.raises(built.PairingError, match="group order"):
        built.Pairing(big_r)
    small_q = PARAMS["a"].replace("q %d" % A["q"], "q %d" % (A["q"] >> 16))
    with pytest.raises(built.PairingError, match="505..512"):
        built.Pairing(small_q)
    with pytest.raises(built.PairingError, match="exp1/exp2"):
        built.Pairing(PARAMS["a"].replace("exp2 159", "exp2 100000"))
    for name in ("f", "d159"):
        T = synth.parse_param(PARAMS[name])
        with pytest.raises(built.PairingError, match="153..159"):
            built.Pairing(PARAMS[name].replace("q %d" % T["q"], "q %d" % (T["q"] >> 8)))
    # scalars travel with the width the kernels stride by
    for name in ("a", "f", "d159", "g149"):
        p = built.Pairing(PARAMS[name])
        r = synth.parse_param(PARAMS[name])["r"]
        assert p.zr_len == (r.bit_length() + 7) // 8


def test_python_mirror_checks_buffer_lengths(built):
    from pbc_b200.params import PARAMS
    p = built.Pairing(PARAMS["a"])
    with pytest.raises(ValueError):
        p.prod_apply(b"\0" * 128, b"\0" * 128, 0, 1)
    with pytest.raises(ValueError):
        p.prod_apply(b"\0" * 128, b"\0" * 256, 2, 1)
    with pytest.raises(ValueError):
        p.pp_apply(b"\0" * 127, b"\0" * 128, 1)
    with pytest.raises(ValueError):
        p.g1_pow_zn(b"\0" * 256, b"\0" * 20, 2)
    with pytest.raises(ValueError):
        p.gt_pow_zn(b"\0" * 128, b"\0" * 19, 1)
    with pytest.raises(ValueError):
        p.g1_from_hash(b"abc", 2, 2)


def test_param_text_tolerates_comments_and_whitespace(built):
    from pbc_b200.params import PARAMS
    text = "# leading comment\n\n" + PARAMS["a"].replace("\n", "   # trailing\n\t")
    p = built.Pairing(text)
    assert p.gt_len == 128


def test_no_gpu_means_loud_failure(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pbc_b200.params import PARAMS
    p = built.Pairing(PARAMS["a"])
    with pytest.raises(built.PairingError, match="CUDA"):
        p.apply(b"\0" * 128, b"\0" * 128, 1)


def test_init_types_f_and_d_lengths(built):
    from pbc_b200.params import PARAMS
    f = built.Pairing(PARAMS["f"])
    assert (f.type, f.g1_len, f.g2_len, f.gt_len) == ("f", 40, 80, 240)
    d = built.Pairing(PARAMS["d159"])
    assert (d.type, d.g1_len, d.g2_len, d.gt_len) == ("d", 40, 120, 120)
    with pytest.raises(built.PairingError, match="k = "):
        built.Pairing(PARAMS["d159"].replace("\nk 6\n", "\nk 4\n"))


def test_derived_constants_match_oracle(built):
    """f_init_pairing / d_init_pairing constants (ecc/f_param.c:408-444, ecc/d_param.c:1035-1049)
    derived by the library's own host big-integer code vs the oracle's restatement."""
    from pbc_b200.params import PARAMS
    from oracle import pbc_oracle as O
    f, of = built.Pairing(PARAMS["f"]), O.pairing_from_param(PARAMS["f"])
    assert tuple(f.derived_constant("xi", 20)) == of.negalpha
    assert tuple(f.derived_constant("xi_inv", 20)) == of.negalphainv
    assert tuple(f.derived_constant("twist_b", 20)) == of.Et.b
    assert tuple(f.derived_constant("xpowq2", 20)) == of.xpowq2
    assert tuple(f.derived_constant("xpowq6", 20)) == of.xpowq6
    assert tuple(f.derived_constant("xpowq8", 20)) == of.xpowq8
    assert f.derived_constant("tateexp", 64) == [of.tateexp]
    d, od = built.Pairing(PARAMS["d159"]), O.pairing_from_param(PARAMS["d159"])
    x = (0, 1, 0)
    x3 = od.Fq3.mul(od.Fq3.mul(x, x), x)
    assert tuple(d.derived_constant("xpwr3", 20)) == x3
    assert tuple(d.derived_constant("xpwr4", 20)) == od.Fq3.mul(x3, x)
    assert tuple(d.derived_constant("xpowq", 20)) == od.xpowq
    assert tuple(d.derived_constant("xpowq2", 20)) == od.xpowq2
    assert d.derived_constant("nqrinv", 20) == [od.nqrinv[0]]
    assert d.derived_constant("nqrinv2", 20) == [od.nqrinv2[0]]
    assert d.derived_constant("phikonr", 32) == [od.phikonr]


def test_bn_parameter_and_frobenius_tables(built):
    """Type F init recovers the BN parameter u from q and r and tabulates xi^(i (q^k - 1)/6) in
    the basis the kernels use: the reference's (forced by the b200_reference_basis test switch) or
    the internal one K[z]/(z^6 - xi')."""
    from pbc_b200.params import PARAMS
    from oracle import pbc_oracle as O
    of = O.pairing_from_param(PARAMS["f"])
    q, F2 = of.q, of.Fq2
    f = built.Pairing(PARAMS["f"] + "b200_reference_basis 1\n")
    (u,) = f.derived_constant("bn_u", 8)
    assert 36 * u ** 4 + 36 * u ** 3 + 24 * u ** 2 + 6 * u + 1 == q
    for k in (1, 2, 3):
        g = F2.pow(of.negalpha, (q ** k - 1) // 6)
        want = []
        for i in range(1, 6):
            want.extend(F2.pow(g, i))
        assert f.derived_constant("frob%d" % k, 20) == want
    assert tuple(f.derived_constant("frob2", 20)[:2]) == of.xpowq2
    with pytest.raises(built.PairingError):
        f.derived_constant("basis_tau", 20)


def test_internal_basis_constants_match_prototype(built):
    """sigma, xi', tau of the isomorphic tower (engine.cu find_f_basis) vs the executable
    specification tools/proto_f_nice_basis.py; tau^6 xi' = phi2(xi)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from proto_f_nice_basis import find_basis
    from pbc_b200.params import PARAMS
    from oracle import pbc_oracle as O
    of = O.pairing_from_param(PARAMS["f"])
    q = of.q
    B = find_basis(q, of.Fq2.nqr, of.negalpha)
    f = built.Pairing(PARAMS["f"])
    assert f.derived_constant("basis_sigma", 20) == [B["sigma"]]
    assert tuple(f.derived_constant("basis_xi_small", 4)) == B["xi_small"]
    assert tuple(f.derived_constant("basis_tau", 20)) == B["tau"]
    K = O.QuadExt(of.Fq, q - 1)
    assert K.mul(K.pow(B["tau"], 6), B["xi_small"]) == B["xi1"]
    for k in (1, 2, 3):
        g = K.pow(B["xi_small"], (q ** k - 1) // 6)
        want = []
        for i in range(1, 6):
            want.extend(K.pow(g, i))
        assert f.derived_constant("frob%d" % k, 20) == want


def test_internal_cubic_basis_constants_match_prototype(built):
    """s, lam, p of the Type D internal cubic w^3 + p w + 1 (engine.cu init_type_d) vs the executable
    specification tools/proto_d_basis.py; the Frobenius constant x^q is re-derived in that basis."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from proto_d_basis import find_basis
    from pbc_b200.params import PARAMS
    from oracle import pbc_oracle as O
    od = O.pairing_from_param(PARAMS["d159"])
    q = od.q
    B = find_basis(q, *od.Fq3.low)
    d = built.Pairing(PARAMS["d159"])
    assert d.derived_constant("basis_cubic", 20) == [B["s"], B["lam"], B["p"]]
    F3n = O.PolyModExt(od.Fq, [1, B["p"], 0])
    assert tuple(d.derived_constant("xpowq_in_use", 20)) == F3n.pow((0, 1, 0), q)
    ref = built.Pairing(PARAMS["d159"] + "b200_reference_basis 1\n")
    assert tuple(ref.derived_constant("xpowq_in_use", 20)) == od.xpowq
    with pytest.raises(built.PairingError):
        ref.derived_constant("basis_cubic", 20)


def test_type_g_init_and_constants(built):
    """g_init_pairing (ecc/g_param.c:1248-1353): sizes, the reduction rows x^5..x^8, the Frobenius
    constants x^q..x^(4q) and (q^4 - q^3 + q^2 - q + 1)/r vs the oracle."""
    from pbc_b200.params import PARAMS
    from oracle import pbc_oracle as O
    g, og = built.Pairing(PARAMS["g149"]), O.pairing_from_param(PARAMS["g149"])
    assert (g.type, g.g1_len, g.g2_len, g.gt_len, g.zr_len) == ("g", 38, 190, 190, 19)
    F5 = og.Fq5
    x = (0, 1, 0, 0, 0)
    xp = F5.pow(x, 5)
    want = []
    for _ in range(4):
        want.extend(xp)
        xp = F5.mul(xp, x)
    assert g.derived_constant_n("xpwr", 19, 20) == want
    want = []
    for i in range(1, 5):
        want.extend(og.xpowq[i])
    assert g.derived_constant_n("xpowq", 19, 20) == want
    assert g.derived_constant("phikonr", 64) == [og.phikonr]
    with pytest.raises(built.PairingError, match="k = "):
        built.Pairing(PARAMS["g149"].replace("\nk 10\n", "\nk 12\n"))


def test_type_a1_parameters_are_accepted_and_checked(built):
    """a1_init_pairing (ecc/a_param.c:2230-2273): p, n, l with p = l n - 1; lengths follow p"""
    from pbc_b200.params import PARAMS
    p = built.Pairing(PARAMS["a1"])
    assert (p.type, p.g1_len, p.g2_len, p.gt_len, p.zr_len) == ("a1", 260, 260, 260, 128)
    with pytest.raises(built.PairingError, match="l\\*n != p\\+1"):
        built.Pairing(PARAMS["a1"].replace("l 1340", "l 1344"))
    with pytest.raises(built.PairingError, match="missing param"):
        built.Pairing("\n".join(l for l in PARAMS["a1"].splitlines() if not l.startswith("n ")))
    small = "type a1\np 153016138656427115508745024995162897686141683994003\nn 1416816098670621439895787268473730534130941518463\nl 108\n"
    q = built.Pairing(small)
    assert (q.g1_len, q.gt_len, q.zr_len) == (42, 42, 20)
