"""The oracle against its own committed golden vectors (tests/golden/pbs_small.npz, made by
tests/golden/make_golden.py).  The GPU tests compare the engine with the same file."""
import os

import numpy as np

from conftest import ROOT, monomial_mul

G = np.load(os.path.join(ROOT, "tests", "golden", "pbs_small.npz"))


def test_oracle_reproduces_golden(small_oracle):
    o, keys = small_oracle
    assert int(G["n"]) == o.n and int(G["keygen_seed"]) == 11
    assert np.bitwise_xor.reduce(keys.ksk.ravel()) == G["ksk_checksum"]
    assert np.bitwise_xor.reduce(keys.bsk.ravel()) == G["bsk_checksum"]
    assert np.array_equal(o.encrypt_big(keys, G["vals"], seed=5), G["cts"])
    assert np.array_equal(o.keyswitch(keys, G["cts"]), G["ks"])
    assert np.array_equal(o.lut_poly(G["table"]), G["lut"])
    for b, e in enumerate(G["cmux_e"]):
        with np.errstate(over="ignore"):
            diff = monomial_mul(G["cmux_glwe"][b], int(e)) - G["cmux_glwe"][b]
        assert np.array_equal(o.external_product_exact(keys.bsk[0], diff, G["cmux_glwe"][b]), G["cmux_exact"][b])
    # value 21 = 16 + 5 has the padding bit set: -f(5)
    table = G["table"].astype(np.int64)
    want = np.array([table[v] if v < 16 else (-table[v - 16]) % 16 for v in G["vals"]])
    assert np.array_equal(G["pbs_exact_decrypt"], want)
