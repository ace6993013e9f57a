"""Generate tests/golden/pbs_small.npz with the oracle (run here, commit the .npz).

The reference holds no golden ciphertexts, so these vectors pin the ORACLE: a later
change to oracle/tfhe_oracle.c that alters a keyswitch output word, a LUT polynomial or an exact
accumulator shows up as a fixture mismatch, and the GPU tests compare the engine against the same
frozen numbers.  Parameters: PARAM_MESSAGE_2_CARRY_2_KS_PBS with n cut to 8 (N = 2048 kept).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import monomial_mul  # noqa: E402


def main():
    p = dict(P)
    p.update(n=8)
    o = Oracle(**p)
    keys = o.keygen(11)
    vals = np.array([0, 3, 7, 12, 15, 21], np.int64)  # 21 has the padding bit set
    cts = o.encrypt_big(keys, vals, seed=5)
    ks = o.keyswitch(keys, cts)
    table = np.array([(3 * x + 1) % 16 for x in range(16)], np.uint8)
    lut = o.lut_poly(table)
    rng = np.random.default_rng(3)
    glwe = rng.integers(0, 2**64, (2, 2, 2048), dtype=np.uint64)
    es = np.array([777, 2048 + 5])
    cmux = []
    for b, e in enumerate(es):
        with np.errstate(over="ignore"):
            diff = monomial_mul(glwe[b], int(e)) - glwe[b]
        cmux.append(o.external_product_exact(keys.bsk[0], diff, glwe[b]))
    pbs = o.pbs_exact(keys, lut[None], [0] * len(vals), cts)
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbs_small.npz"),
        keygen_seed=11, n=8, vals=vals, cts=cts, ks=ks, table=table, lut=lut,
        cmux_glwe=glwe, cmux_e=es, cmux_exact=np.stack(cmux),
        pbs_exact_decrypt=o.decrypt_big(keys, pbs),
        ksk_checksum=np.bitwise_xor.reduce(keys.ksk.ravel()), bsk_checksum=np.bitwise_xor.reduce(keys.bsk.ravel()),
    )
    print("wrote pbs_small.npz")


if __name__ == "__main__":
    main()
