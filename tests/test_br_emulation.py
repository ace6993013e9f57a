"""The blind-rotation thread program (fhestring_b200/csrc/br_core.cuh) is compiled for the host and run
by 64 std::threads (tests/emu/br_emu.cpp): index logic, four-step FFT layout, Fourier-BSK layout,
mod-switch, rotation and sample extract are checked against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, monomial_mul


# every build of the thread program: the default kernel and the settings of its two compile-time knobs (br_core.cuh)
VARIANTS = {
    "default": (),                                        # no torus conversion on the FP64 pipe, 6 key rows prefetched
    "fp64_conversions": ("FHESTR_BR_CVT_FP64=1",),        # every torus conversion by the 1.5 * 2^52 trick
    "every_4th_fp64_depth_12": ("FHESTR_BR_CVT_FP64=4", "FHESTR_BR_PREFETCH=12"),
}


@pytest.fixture(scope="module", params=list(VARIANTS), ids=list(VARIANTS))
def emu(request):
    tag = request.param
    src = os.path.join(ROOT, "tests", "emu", "br_emu.cpp")
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, f"libbr_emu_{tag}.so")
    deps = [src] + [os.path.join(ROOT, "fhestring_b200", "csrc", f) for f in ("br_core.cuh", "fft32_gen.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-shared", "-fPIC",
                               *[f"-D{d}" for d in VARIANTS[tag]], "-o", lib, src])
    return C.CDLL(lib)


def _p(a, t=C.c_uint64):
    return a.ctypes.data_as(C.POINTER(t))


def _convert(emu, bsk):
    n = bsk.shape[0]
    out = np.zeros((n, 4096, 2), np.float64)
    emu.emu_convert_bsk(n, _p(np.ascontiguousarray(bsk)), _p(out, C.c_double))
    return out


def test_single_cmux_against_exact(emu, small_oracle):
    o, keys = small_oracle
    bskf = _convert(emu, keys.bsk[:1])
    rng = np.random.default_rng(1)
    for e in (1, 1234, 2048, 2048 + 77, 4095):
        # accumulator resolution of the kernel: 32-bit torus (br_core.cuh acc_t)
        glwe = rng.integers(0, 2**64, (2, 2048), dtype=np.uint64) & np.uint64(0xFFFFFFFF00000000)
        ks = np.array([np.uint64(e) << np.uint64(52), 0], np.uint64)
        got = np.zeros((2, 2048), np.uint64)
        emu.emu_blind_rotate(1, _p(ks), None, _p(glwe), _p(bskf, C.c_double), None, _p(got))
        with np.errstate(over="ignore"):
            diff = monomial_mul(glwe, e) - glwe
        want = o.external_product_exact(keys.bsk[0], diff, glwe)
        d = (got - want).astype(np.int64).astype(float)
        assert np.sqrt(np.mean(d * d)) < 2.0**40, e   # 2^-24 of the torus
        assert np.abs(d).max() < 2.0**43, e


def test_zero_rotation_is_skipped_exactly(emu, small_oracle):
    o, keys = small_oracle
    bskf = _convert(emu, keys.bsk[:1])
    glwe = np.random.default_rng(2).integers(0, 2**64, (2, 2048), dtype=np.uint64)
    ks = np.zeros(2, np.uint64)
    got = np.zeros((2, 2048), np.uint64)
    emu.emu_blind_rotate(1, _p(ks), None, _p(glwe), _p(bskf, C.c_double), None, _p(got))
    # the input rounded to the 32-bit torus, nothing else
    assert np.array_equal(got, ((glwe + np.uint64(1 << 31)) >> np.uint64(32)) << np.uint64(32))


def test_small_pbs_decrypts(emu, small_oracle):
    o, keys = small_oracle
    n = 3
    bskf = _convert(emu, keys.bsk[:n])
    # a 3-step rotation is a valid PBS for the sub-key (s_lwe[:3]) -- re-keyswitch with that key
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    p = dict(P); p.update(n=n)
    o3 = Oracle(**p)
    k3 = o3.keygen(5)
    bskf = _convert(emu, k3.bsk)
    table = [(3 * x + 1) % 16 for x in range(16)]
    lut = o3.lut_poly(table)
    vals = np.array([0, 5, 9, 15])
    cts = o3.encrypt_big(k3, vals, seed=2)
    ks = o3.keyswitch(k3, cts)
    outs = np.zeros((len(vals), 2049), np.uint64)
    for b in range(len(vals)):
        emu.emu_blind_rotate(n, _p(ks[b]), _p(lut), None, _p(bskf, C.c_double), _p(outs[b]), None)
    assert np.array_equal(o3.decrypt_big(k3, outs), np.array([table[v] for v in vals]))
