"""Ciphertext-level parity with tfhe-rs 0.5.2, from golden vectors the Rust harness `integration/fhestr-parity` writes.

The build image has no Rust toolchain, so the real file (tests/golden/tfhe_rs_fixture.bin) cannot be produced here:
those tests SKIP until a maintainer runs `cargo run --release -- fixture ../../tests/golden/tfhe_rs_fixture.bin` on a
box with cargo -- from then on they are the pass/fail for "bit-exact keyswitch / LUT / sample extract against tfhe-rs"
(SURVEY.md 8c(3)).  The SAME checks always run on a synthetic file of the same layout written from the oracle's own
keys, so the reader, the check code and the engine path are exercised in every round."""
import os

import numpy as np
import pytest

from conftest import ROOT
from fhestring_b200 import fixtures

REAL = os.path.join(ROOT, "tests", "golden", "tfhe_rs_fixture.bin")


def synthetic(tmp_path_factory):
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    p = dict(P); p.update(n=8)
    o = Oracle(**p)
    keys = o.keygen(77)
    tables = np.array([list(range(16)), [int((x >> 2) == (x & 3)) for x in range(16)], [(3 * x + 1) % 16 for x in range(16)],
                       [int(x != 0) for x in range(16)]], np.uint8)
    bodies = np.stack([o.lut_poly(t) for t in tables])
    vals = (np.arange(24) * 7 + 3) % 32
    cts = o.encrypt_big(keys, vals, seed=78)
    lut = (np.arange(24) % len(tables)).astype(np.uint32)
    fx = fixtures.Fixture(
        params={k: int(p[k]) for k in fixtures.PARAM_FIELDS}, bsk_std=keys.bsk, ksk=keys.ksk, s_lwe=keys.s_lwe, s_glwe=keys.s_glwe,
        tables=tables, lut_bodies=bodies, triple_lut=lut, triple_in=cts, triple_ks=o.keyswitch(keys, cts),
        triple_out=o.pbs_fft(keys, o.fourier_bsk(keys), bodies, lut.astype(np.int32), cts)[0])
    path = str(tmp_path_factory.mktemp("fx") / "synthetic.bin")
    fixtures.save(path, fx)
    return path


@pytest.fixture(scope="module", params=["synthetic", "tfhe_rs"])
def fx(request, tmp_path_factory):
    if request.param == "tfhe_rs":
        if not os.path.exists(REAL):
            pytest.skip("tests/golden/tfhe_rs_fixture.bin absent: produce it with integration/fhestr-parity on a box with cargo")
        return fixtures.load(REAL)
    return fixtures.load(synthetic(tmp_path_factory))


def _oracle(fx):
    from oracle.tfhe_oracle import Keys, Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    p = dict(P); p.update({k: v for k, v in fx.params.items() if k in P})
    o = Oracle(**p)
    return o, Keys(s_lwe=fx.s_lwe, s_glwe=fx.s_glwe, bsk=np.ascontiguousarray(fx.bsk_std), ksk=np.ascontiguousarray(fx.ksk))


def _expected(fx):
    v = (np.arange(len(fx.triple_lut)) * 7 + 3) % 32          # the harness' plaintexts (main.rs: value = (7 i + 3) % 32)
    f = fx.tables[fx.triple_lut, v % 16].astype(np.int64)
    return np.where(v < 16, f, (-f) % 16)


def test_oracle_keyswitch_lut_and_pbs_against_fixture(fx):
    """CPU: the oracle against the file -- keyswitch words and LUT bodies bit-exact, PBS outputs decrypt identically and
    sit within the scheme's rounding noise of the file's phases"""
    o, keys = _oracle(fx)
    assert np.array_equal(o.keyswitch(keys, fx.triple_in), fx.triple_ks)                       # K0 + K1
    for t, body in zip(fx.tables, fx.lut_bodies):
        assert np.array_equal(o.lut_poly(t), body)                                             # K5
    want = _expected(fx)
    assert np.array_equal(o.decrypt_big(keys, fx.triple_out), want)                            # the file is self-consistent
    out = o.pbs_exact(keys, fx.lut_bodies, fx.triple_lut.astype(np.int32), fx.triple_in)      # K2..K4, exact integers
    assert np.array_equal(o.decrypt_big(keys, out), want)
    dp = (o.phases(keys.s_glwe, out) - o.phases(keys.s_glwe, fx.triple_out)).astype(np.int64).astype(float) / 2.0**64
    assert np.abs(dp).max() < 2.0**-11


@pytest.mark.gpu
def test_engine_against_fixture(fx, build_lib):
    """B200: the engine on the file's keys -- keyswitch and LUT words bit-exact, PBS decrypts identically"""
    from fhestring_b200.engine import Engine, single_term_jobs
    o, keys = _oracle(fx)
    eng = Engine(arena_blocks=1024, **{k: v for k, v in fx.params.items()})
    try:
        eng.load_keys(fx.bsk_std, fx.ksk)
        T = len(fx.triple_lut)
        eng.upload(0, fx.triple_in)
        ids = [eng.lut(t) for t in fx.tables]
        for i, body in zip(ids, fx.lut_bodies):
            assert np.array_equal(eng.lut_download(i), body)
        jobs = single_term_jobs(512 + np.arange(T), np.arange(T), 0)
        jobs["lut"] = [ids[int(l)] for l in fx.triple_lut]
        assert np.array_equal(eng.debug_keyswitch(jobs), fx.triple_ks)
        for mode in (1, 3, 4):
            eng.set_br_mode(mode)
            eng.pbs_batch(jobs)
            got = eng.download(512, T)
            assert np.array_equal(o.decrypt_big(keys, got), _expected(fx)), mode
            dp = (o.phases(keys.s_glwe, got) - o.phases(keys.s_glwe, fx.triple_out)).astype(np.int64).astype(float) / 2.0**64
            assert np.abs(dp).max() < 2.0**-11
    finally:
        eng.close()
