"""The oracle (oracle/pbc_oracle.py) pinned against the reference's own known-answer vectors
and against fixtures produced by the compiled, unmodified reference (tests/golden/)."""
import pytest

from oracle import pbc_oracle as O
from pbc_b200.params import PARAMS

# pbc/pairing_test.pbc:3-21 (Type A, a.param) -- the only fixed outputs the reference ships.
KAT_G = (2382389466570123849673299401984867521337122094157231907755149435707124249269394670242462497382963719723036281844079382411446883273020125104982896098602669,
         2152768906589770702756591740710760107878949212304343787392475836859241438597588807103470081101790991563152395601123682809718038151417122294066319979967168)
KAT_H = (5832612417453786541700129157230442590988122495898645678468800815872828277169950107203266157735206975228912899931278160262081308603240860553459187732968543,
         5825590786822892934138376868455818413990615826926356662470129700411774690868351658310187202553513693344017463065909279569624651155563430675084173630054336)
KAT_A = 171583727262251826931173602797951212789946235851
KAT_B = 233634857565210859330459959563397971304462340857
KAT_E_GH = (1352478452661998164151215014828915385601138645645403926287105573769451214277485326392786454433874957123922454604362337349978217917242114505658729401276644,
            2809858014072341042857607405424304552357466023841122154308055820747972163307396014445308786731013691659356362568425895483877936945589613445697089590886519)
KAT_GA = (3727290142167731134589933003026410141163353118002821914170365887139605219852868537686435214464927363733592858325260588072422405672197113236445369761687270,
          8313413520789037477320458888316489483781506373846006723006557775349684878102042826049292521482530556981023752851151672326421296204733037418468523296005577)
KAT_HB = (302169045606583472168811217560382970305157511680176350745436990853463473855962841196184541109617397027480204774682450915021848512168573082843355648090809,
          7428193877404140917518137438384425427600294220905786853638038223349096573857683866658575603565175187399696035468569929483731011292133989973187846752806084)
KAT_RES = (5401677742232403160612802517983583823254857216272776607059355607024091426935935872461700304196658606704085604766577186374528948004140797833341187234647180,
           4255900207739859478558185000995524505026245539159946661271849714832846423204570340979120001638894488614502770175520505048836617405342161594891740961421000)


@pytest.fixture(scope="module")
def pa():
    return O.pairing_from_param(PARAMS["a"])


def test_kat_type_a_pairing(pa):
    # pbc/pairing_test.pbc:11
    assert pa.pairing(KAT_G, KAT_H) == KAT_E_GH


def test_kat_type_a_scalar_mults_and_bilinearity(pa):
    # pbc/pairing_test.pbc:13-21
    ga = pa.E.mul(KAT_A, KAT_G)
    hb = pa.E.mul(KAT_B, KAT_H)
    assert ga == KAT_GA and hb == KAT_HB
    assert pa.pairing(ga, hb) == KAT_RES
    assert pa.Fq2.pow(KAT_E_GH, (KAT_A * KAT_B) % pa.r) == KAT_RES
    assert pa.Fq2.pow(pa.pairing(ga, KAT_H), KAT_B) == KAT_RES


def test_kat_prod_and_pp_agree(pa):
    # benchmark/benchmark.c:93-96 self-check: pp_apply == element_pairing
    assert pa.pp_pairing(KAT_G, KAT_H) == KAT_E_GH
    e2 = pa.prod_pairing([KAT_G, KAT_GA], [KAT_H, KAT_HB])
    assert e2 == pa.Fq2.mul(KAT_E_GH, KAT_RES)


@pytest.mark.parametrize("name,limit", [("a", 24), ("d159", 8), ("f", 3), ("g149", 4), ("a1", 3)])
def test_oracle_matches_reference_fixtures(golden, name, limit):
    g = golden[name]
    pr = O.pairing_from_param(PARAMS[name])
    assert (pr.g1_len, pr.g2_len, pr.gt_len, pr.zr_len) == tuple(g["lengths"][k] for k in ("g1", "g2", "gt", "zr"))
    for P, Q, e in list(zip(g["pairing"]["P"], g["pairing"]["Q"], g["pairing"]["e"]))[:limit]:
        assert O.pairing_bytes(pr, bytes.fromhex(P), bytes.fromhex(Q)).hex() == e


@pytest.mark.parametrize("name,limit", [("a", 3), ("d159", 2), ("f", 1), ("a1", 1)])
def test_oracle_prod_matches_reference_fixtures(golden, name, limit):
    g = golden[name]["prod"]
    pr = O.pairing_from_param(PARAMS[name])
    k = g["k"]
    for i, e in enumerate(g["e"][:limit]):
        Ps = [bytes.fromhex(x) for x in g["P"][i * k:(i + 1) * k]]
        Qs = [bytes.fromhex(x) for x in g["Q"][i * k:(i + 1) * k]]
        assert O.prod_pairing_bytes(pr, Ps, Qs).hex() == e


@pytest.mark.parametrize("name", ["a", "d159", "f", "a1"])
def test_oracle_offcurve_is_identity(golden, name):
    g = golden[name]
    pr = O.pairing_from_param(PARAMS[name])
    P0 = bytes.fromhex(g["pairing"]["P"][0])
    Q0 = bytes.fromhex(g["pairing"]["Q"][0])
    ident = g["offcurve"]["identity"]
    assert O.pairing_bytes(pr, bytes.fromhex(g["offcurve"]["badP"]), Q0).hex() == ident
    assert O.pairing_bytes(pr, P0, bytes.fromhex(g["offcurve"]["badQ"])).hex() == ident
    assert pr.GT.to_bytes(pr.GT.one).hex() == ident


def test_pp_fixture_equals_plain_pairing(golden):
    for name in ("a", "d159", "f", "a1"):
        g = golden[name]
        assert g["pp"]["e"] == g["pairing"]["e"][:1] + g["pp"]["e"][1:]  # first Q pairs with P[0]


def test_param_parser_rejects_garbage():
    with pytest.raises(ValueError):
        O.pairing_from_param("q 17\nr 3\n")
    with pytest.raises(ValueError):
        O.pairing_from_param("type zz\nq 17\n")


def test_from_hash_restatement_matches_reference_fixtures(golden):
    """curve_from_hash (ecc/curve.c:455-482) + pbc_mpz_from_hash (arith/field.c:643-668): the
    oracle's restatement vs G1 elements the compiled reference derived from the same bytes."""
    from oracle import pbc_oracle as O
    from pbc_b200.params import PARAMS
    for name in ("a", "f", "d159", "g149", "a1"):
        orc = O.pairing_from_param(PARAMS[name])
        for ln, blk in golden[name]["hash"].items():
            for d, want in zip(blk["data"], blk["G1"]):
                assert len(d) == 2 * int(ln)
                assert orc.G1.to_bytes(O.g1_from_hash(orc, bytes.fromhex(d))).hex() == want


def test_type_a1_pp_restatement_matches_reference_fixture(golden):
    """a1_pairing_pp_init / _apply (ecc/a_param.c:1632-1818, tangent and chord merged into one
    conic per set bit of n) restated in the oracle vs the compiled reference's pp outputs."""
    g = golden["a1"]
    pr = O.pairing_from_param(PARAMS["a1"])
    P0 = pr.G1.from_bytes(bytes.fromhex(g["pp"]["P"]))
    for Q, e in list(zip(g["pairing"]["Q"], g["pp"]["e"]))[:2]:
        assert pr.GT.to_bytes(pr.pp_pairing(P0, pr.G2.from_bytes(bytes.fromhex(Q)))).hex() == e
