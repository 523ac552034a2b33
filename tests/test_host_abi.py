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
