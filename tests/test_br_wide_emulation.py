"""The latency form of the blind rotation (fhestring_b200/csrc/br_wide.cuh: one PBS over 128 threads, both GLWE
polynomials in every thread, three radix-8 stages + a half level, Fourier key in its own spectrum order) is compiled
for the host and run by 128 std::threads (tests/emu/br_wide_emu.cpp): transform pair, key layout, rotation,
mod-switch and sample extract are checked against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, monomial_mul


# both forms of the CMUX step: the plain one (one barrier for both polynomials: the pair kernel) and the two-stream
# software pipeline with one barrier per polynomial (the single kernel)
@pytest.fixture(scope="module", params=[0, 1], ids=["plain_step", "pipelined_step"])
def emu(request):
    src = os.path.join(ROOT, "tests", "emu", "br_wide_emu.cpp")
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"libbr_wide_emu_{request.param}.so")
    deps = [src] + [os.path.join(ROOT, "fhestring_b200", "csrc", f) for f in ("br_wide.cuh", "br_core.cuh", "fft32_gen.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC", f"-DEMU_PIPELINED={request.param}",
                               "-o", lib, src])
    return C.CDLL(lib)


def _p(a, t=C.c_uint64):
    return a.ctypes.data_as(C.POINTER(t))


def _convert(emu, bsk):
    n = bsk.shape[0]
    out = np.zeros((n, 4096, 2), np.float64)
    emu.emu_wide_convert_bsk(n, _p(np.ascontiguousarray(bsk)), _p(out, C.c_double))
    return out


def test_transform_pair_and_spectrum(emu):
    """inverse(forward(c)) = 1024 c, and the spectrum is a permutation of c evaluated at the roots of x^1024 = i"""
    rng = np.random.default_rng(0)
    re, im = rng.standard_normal(1024), rng.standard_normal(1024)
    ore, oim, spec = np.zeros(1024), np.zeros(1024), np.zeros((8, 128, 2))
    emu.emu_wide_roundtrip(_p(re, C.c_double), _p(im, C.c_double), _p(ore, C.c_double), _p(oim, C.c_double), _p(spec, C.c_double))
    assert np.abs(ore / 1024 - re).max() < 1e-13 and np.abs(oim / 1024 - im).max() < 1e-13
    roots = np.exp(1j * np.pi * (1 + 4 * np.arange(1024)) / 2048)
    vals = np.polyval((re + 1j * im)[::-1], roots)
    s = (spec[..., 0] + 1j * spec[..., 1]).ravel()
    d = np.abs(s[:, None] - vals[None, :])
    assert (d.min(axis=1) < 1e-9).all() and len(set(d.argmin(axis=1))) == 1024


def test_single_cmux_against_exact(emu, small_oracle):
    o, keys = small_oracle
    bskw = _convert(emu, keys.bsk[:1])
    rng = np.random.default_rng(1)
    for e in (0, 1, 1234, 2048, 2048 + 77, 4095):
        glwe = rng.integers(0, 2**64, (2, 2048), dtype=np.uint64) & np.uint64(0xFFFFFFFF00000000)
        ks = np.array([np.uint64(e) << np.uint64(52), 0], np.uint64)
        got = np.zeros((2, 2048), np.uint64)
        emu.emu_wide_blind_rotate(1, _p(ks), None, _p(glwe), _p(bskw, C.c_double), None, _p(got))
        with np.errstate(over="ignore"):
            diff = monomial_mul(glwe, e) - glwe
        want = o.external_product_exact(keys.bsk[0], diff, glwe)
        d = (got - want).astype(np.int64).astype(float)
        assert np.sqrt(np.mean(d * d)) < 2.0**40, e   # 2^-24 of the torus
        assert np.abs(d).max() < 2.0**43, e
        if e == 0:
            assert np.array_equal(got, glwe)          # nothing to add: exact


def test_small_pbs_decrypts(emu):
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    n = 3
    p = dict(P); p.update(n=n)
    o3 = Oracle(**p)
    k3 = o3.keygen(5)
    bskw = _convert(emu, k3.bsk)
    table = [(3 * x + 1) % 16 for x in range(16)]
    lut = o3.lut_poly(table)
    vals = np.array([0, 5, 9, 15])
    cts = o3.encrypt_big(k3, vals, seed=2)
    ks = o3.keyswitch(k3, cts)
    outs = np.zeros((len(vals), 2049), np.uint64)
    for b in range(len(vals)):
        emu.emu_wide_blind_rotate(n, _p(ks[b]), _p(lut), None, _p(bskw, C.c_double), _p(outs[b]), None)
    assert np.array_equal(o3.decrypt_big(k3, outs), np.array([table[v] for v in vals]))


def test_pipelined_step_is_bit_identical_to_the_plain_step(small_oracle):
    """same arithmetic in the same order per polynomial: the two forms of the step must produce the same words"""
    o, keys = small_oracle
    outs = []
    for flag in (0, 1):
        lib = os.path.join(ROOT, "tests", "_build", f"libbr_wide_emu_{flag}.so")
        if not os.path.exists(lib):
            pytest.skip("emulation libraries not built (run the whole module)")
        e = C.CDLL(lib)
        bskw = _convert(e, keys.bsk[:3])
        glwe = np.random.default_rng(9).integers(0, 2**64, (2, 2048), dtype=np.uint64)
        ks = np.array([np.uint64(1234) << np.uint64(52), np.uint64(77) << np.uint64(52), np.uint64(4000) << np.uint64(52), 0], np.uint64)
        got = np.zeros((2, 2048), np.uint64)
        e.emu_wide_blind_rotate(3, _p(ks), None, _p(glwe), _p(bskw, C.c_double), None, _p(got))
        outs.append(got)
    assert np.array_equal(outs[0], outs[1])

