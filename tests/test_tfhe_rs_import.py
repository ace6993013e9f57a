"""fhestring_b200/tfhe_rs_import.py: the bincode reader of a serialised tfhe::integer::ServerKey, against a writer of the
same (recalled, unverified) layout: round trip of a real-sized key set, and loud failures on every structural
mismatch.  No tfhe-rs-written file exists in this image (no Rust toolchain) -- see the module's STATUS note."""
import numpy as np
import pytest

from fhestring_b200 import tfhe_rs_import as ti


def test_round_trip_and_the_recovered_key_bootstraps(small_oracle):
    o, keys = small_oracle
    params = dict(n=8, N=2048, k=1, pbs_base_log=23, pbs_level=1, ks_base_log=3, ks_level=5, delta_log=59)
    raw = ti.write_server_key(params, keys.bsk, keys.ksk)
    p2, bsk, ksk = ti.read_server_key(raw)
    assert p2 == params
    assert np.array_equal(ksk, keys.ksk)
    # the Fourier round trip is exact for these magnitudes up to f64 rounding: at most a few hundred ulps of 2^-64
    d = (bsk - keys.bsk).astype(np.int64)
    assert np.abs(d).max() < 1 << 14
    # and the recovered key is a working bootstrapping key
    from oracle.tfhe_oracle import Keys
    k2 = Keys(s_lwe=keys.s_lwe, s_glwe=keys.s_glwe, bsk=bsk, ksk=ksk)
    vals = np.arange(16)
    cts = o.encrypt_big(keys, vals, seed=3)
    table = [(3 * x + 1) % 16 for x in range(16)]
    out, _ = o.pbs_fft(k2, o.fourier_bsk(k2), o.lut_poly(table)[None], [0] * 16, cts)
    assert np.array_equal(o.decrypt_big(keys, out), np.array(table))


def test_structural_mismatches_fail_loudly(small_oracle):
    o, keys = small_oracle
    params = dict(n=8, N=2048, k=1, pbs_base_log=23, pbs_level=1, ks_base_log=3, ks_level=5, delta_log=59)
    raw = ti.write_server_key(params, keys.bsk, keys.ksk)
    with pytest.raises(ti.BincodeError):
        ti.read_server_key(raw + b"\0")                       # trailing bytes
    with pytest.raises((ti.BincodeError, Exception)):
        ti.read_server_key(raw[:-9])                          # truncated
    bad = bytearray(raw)
    off = 8 + keys.ksk.size * 8 + 24 + 24                     # the bootstrapping-key enum tag
    bad[off:off + 4] = (1).to_bytes(4, "little")
    with pytest.raises(ti.BincodeError, match="Classic"):
        ti.read_server_key(bytes(bad))
