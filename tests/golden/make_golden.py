"""Generate tests/golden/*.json from the compiled, unmodified reference (oracle/_ref).

NOTE: the first parameter set generated in a process (a) does not reproduce its random draws from
run to run (the reference seeds part of its state at first use); a.json as committed was kept and
its later additions ("pow": Qa, e_Pa_Qa) were computed by the reference from the stored inputs.


Run HERE (container with /root/reference):  make -C oracle && python tests/golden/make_golden.py
The fixtures are what travels to the GPU box; nothing at test time reads /root/reference.
All byte strings are the reference wire format (element_to_bytes), hex-encoded.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402
from pbc_b200.params import PARAMS  # noqa: E402

SEED = 20260922
N_SINGLE = {"a": 24, "f": 12, "d159": 16, "g149": 12, "a1": 8}
PROD = {"a": (4, 3), "f": (3, 2), "d159": (4, 3), "g149": (3, 2), "a1": (3, 2)}  # (k, n_out)


def chunks(b, n):
    return [b[i:i + n].hex() for i in range(0, len(b), n)]


def main():
    only = sys.argv[1:]          # e.g. `make_golden.py g149` regenerates that file alone
    for name, text in PARAMS.items():
        if only and name not in only:
            continue
        rp = R.RefPairing(text)
        R.RefPairing.seed(SEED)
        n = N_SINGLE[name]
        P = rp.random(R.G1, n)
        Q = rp.random(R.G2, n)
        E = rp.pairing(P, Q, n)
        k, n_out = PROD[name]
        PP = rp.random(R.G1, k * n_out)
        QQ = rp.random(R.G2, k * n_out)
        EP = rp.prod_pairing(PP, QQ, k, n_out)
        # fixed-first-argument (pairing_pp_*) outputs: must equal plain pairings
        EPP = rp.pp_pairing(P[:rp.g1_len], Q, 4)
        # bilinearity material: P^a, Q^b, e^(ab)
        a = rp.random(R.ZR, 4)
        Pa = rp.pow_zn(R.G1, P[:4 * rp.g1_len], a, 4)
        Ea = rp.pairing(Pa, Q[:4 * rp.g2_len], 4)
        Qa = rp.pow_zn(R.G2, Q[:4 * rp.g2_len], a, 4)          # same scalars on G2 (no new random draws)
        Eaa = rp.pairing(Pa, Qa, 4)                            # e(P^a, Q^a) = e(P, Q)^(a^2)
        # off-curve inputs decode to O and pair to the GT identity
        badP = bytes([P[0] ^ 1]) + P[1:rp.g1_len]
        badQ = Q[:rp.g2_len - 1] + bytes([Q[rp.g2_len - 1] ^ 1])
        e_badP = rp.pairing(badP, Q[:rp.g2_len], 1)
        e_badQ = rp.pairing(P[:rp.g1_len], badQ, 1)
        assert rp.is_identity(R.GT, e_badP) and rp.is_identity(R.GT, e_badQ)
        doc = {
            "source": "oracle/_ref/libpbcref.so (unmodified reference, commit cdf8e1b), "
                      "pbc_random_set_deterministic(%d)" % SEED,
            "param": name,
            "lengths": {"g1": rp.g1_len, "g2": rp.g2_len, "gt": rp.gt_len, "zr": rp.zr_len},
            "pairing": {"P": chunks(P, rp.g1_len), "Q": chunks(Q, rp.g2_len),
                        "e": chunks(E, rp.gt_len)},
            "prod": {"k": k, "P": chunks(PP, rp.g1_len), "Q": chunks(QQ, rp.g2_len),
                     "e": chunks(EP, rp.gt_len)},
            "pp": {"P": P[:rp.g1_len].hex(), "e": chunks(EPP, rp.gt_len)},
            "pow": {"a": chunks(a, rp.zr_len), "Pa": chunks(Pa, rp.g1_len),
                    "e_Pa_Q": chunks(Ea, rp.gt_len), "Qa": chunks(Qa, rp.g2_len),
                    "e_Pa_Qa": chunks(Eaa, rp.gt_len)},
            "offcurve": {"badP": badP.hex(), "badQ": badQ.hex(), "identity": e_badP.hex()},
        }
        # element_from_hash on G1 (ecc/curve.c:455-482) for a few input lengths
        import random
        rnd = random.Random(SEED)
        doc["hash"] = {}
        for ln in (20, 32, 70, 3):
            data = [bytes(rnd.randrange(256) for _ in range(ln)) for _ in range(6)]
            H = rp.from_hash(R.G1, b"".join(data), ln, len(data))
            doc["hash"][str(ln)] = {"data": [d.hex() for d in data], "G1": chunks(H, rp.g1_len)}
        # element_to_bytes_compressed / element_from_bytes_compressed (ecc/curve.c:762-813)
        nc = min(8, n)
        comp = rp.compress(R.G1, P[:nc * rp.g1_len], nc)
        clen = len(comp) // nc
        flipped = b"".join(comp[i * clen:(i + 1) * clen - 1] + bytes([comp[(i + 1) * clen - 1] ^ 1])
                           for i in range(nc))
        doc["compressed"] = {"len": clen, "G1": chunks(comp, clen),
                             "G1_flipped_sign": chunks(rp.decompress(R.G1, flipped, nc), rp.g1_len)}
        with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
            json.dump(doc, f, indent=0)
        print(name, "ok")


if __name__ == "__main__":
    main()
