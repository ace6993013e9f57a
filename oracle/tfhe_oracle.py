"""ctypes binding of oracle/tfhe_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (fhestring_b200/) never does.  See the header of tfhe_oracle.c for
what it restates (reference call sites fheasciichar.rs:23-102) and for the
parity status (ciphertext level: UNPINNED -- tfhe-rs is not in the reference tree).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "tfhe_oracle.c")
_LIB = os.path.join(_HERE, "_build", "libtfhe_oracle.so")


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp the oracle into oracle/_build/ (generic x86-64 + runtime-dispatched clones)."""
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O3", "-fno-math-errno", "-fno-trapping-math", "-fopenmp", "-shared", "-fPIC", "-std=gnu11", "-o", _LIB, _SRC, "-lm"]
        )
    return _LIB


class OrcParams(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32),
        ("pbs_base_log", C.c_int32), ("pbs_level", C.c_int32),
        ("ks_base_log", C.c_int32), ("ks_level", C.c_int32),
        ("delta_log", C.c_int32),
        ("lwe_std", C.c_double), ("glwe_std", C.c_double),
    ]


# PARAM_MESSAGE_2_CARRY_2_KS_PBS as recalled in SURVEY.md A.1 (reference: src/main.rs:3,43)
PARAM_MESSAGE_2_CARRY_2_KS_PBS = dict(
    n=742, N=2048, k=1, pbs_base_log=23, pbs_level=1, ks_base_log=3, ks_level=5, delta_log=59,
    lwe_std=7.069849454709433e-6, glwe_std=2.9403601535432533e-16,
)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@dataclass
class Keys:
    s_lwe: np.ndarray
    s_glwe: np.ndarray
    bsk: np.ndarray  # [n][level][2][2][N] u64 standard domain
    ksk: np.ndarray  # [N][ks_level][n+1] u64


class Oracle:
    def __init__(self, **params):
        self.lib = C.CDLL(build())
        self.p = OrcParams(**params)
        L = self.lib
        L.orc_fft_plan_new.restype = C.c_void_p
        L.orc_lwe_phase.restype = C.c_uint64
        L.orc_decode.restype = C.c_uint32
        L.orc_modswitch.restype = C.c_uint32
        L.orc_fourier_bsk_doubles.restype = C.c_size_t
        L.orc_pbs_fft_batch.restype = C.c_int
        L.orc_max_threads.restype = C.c_int

    # ---- sizes
    @property
    def n(self): return self.p.n
    @property
    def N(self): return self.p.N
    @property
    def big(self): return self.p.N * self.p.k + 1
    @property
    def small(self): return self.p.n + 1
    @property
    def delta_log(self): return self.p.delta_log

    def max_threads(self) -> int:
        return int(self.lib.orc_max_threads())

    def set_acc32(self, on: bool):
        """A/B switch: round the f64 route's accumulator to the 32-bit torus where the GPU kernels do (tfhe_oracle.c)"""
        self.lib.orc_set_acc32(C.c_int(1 if on else 0))

    def set_threads(self, n: int) -> int:
        """OpenMP thread count for the batch entry points (torchrun exports OMP_NUM_THREADS=1); returns the count in effect"""
        return int(self.lib.orc_set_threads(C.c_int(int(n))))

    # ---- keys / client side
    def keygen(self, seed: int, want_bsk=True, want_ksk=True) -> Keys:
        p = self.p
        s_lwe = np.zeros(p.n, np.uint8)
        s_glwe = np.zeros(p.N, np.uint8)
        bsk = np.zeros((p.n, p.pbs_level, 2, 2, p.N), np.uint64) if want_bsk else None
        ksk = np.zeros((p.N, p.ks_level, p.n + 1), np.uint64) if want_ksk else None
        self.lib.orc_keygen(C.byref(p), C.c_uint64(seed), _p(s_lwe, C.c_uint8), _p(s_glwe, C.c_uint8),
                            _p(bsk, C.c_uint64) if want_bsk else None,
                            _p(ksk, C.c_uint64) if want_ksk else None)
        return Keys(s_lwe, s_glwe, bsk, ksk)

    def encrypt_big(self, keys: Keys, values, seed: int, std=None) -> np.ndarray:
        """Encrypt block values (0..31 incl. padding bit) under the big key: [count][N+1] u64."""
        values = np.asarray(values, np.uint64).ravel()
        std = self.p.glwe_std if std is None else std
        out = np.zeros((len(values), self.big), np.uint64)
        for i, v in enumerate(values):
            self.lib.orc_lwe_encrypt(_p(keys.s_glwe, C.c_uint8), C.c_int(self.p.N),
                                     C.c_uint64(int(v) << self.p.delta_log), C.c_double(std),
                                     C.c_uint64(seed), C.c_uint64(i), _p(out[i], C.c_uint64))
        return out

    def phases(self, key: np.ndarray, cts: np.ndarray) -> np.ndarray:
        cts = np.ascontiguousarray(cts, np.uint64)
        dim = cts.shape[-1] - 1
        flat = cts.reshape(-1, dim + 1)
        out = np.zeros(flat.shape[0], np.uint64)
        self.lib.orc_lwe_phase_batch(_p(key, C.c_uint8), C.c_int(dim), _p(flat, C.c_uint64),
                                     C.c_int(flat.shape[0]), _p(out, C.c_uint64))
        return out

    def decode(self, phases: np.ndarray) -> np.ndarray:
        """A.2: ((phase + delta/2) >> delta_log) with the padding bit dropped."""
        d = self.p.delta_log
        ph = np.asarray(phases, np.uint64)
        v = (ph + np.uint64(1 << (d - 1))) >> np.uint64(d)
        return (v & np.uint64((1 << (63 - d)) - 1)).astype(np.int64)

    def decrypt_big(self, keys: Keys, cts: np.ndarray) -> np.ndarray:
        return self.decode(self.phases(keys.s_glwe, cts))

    # ---- server side
    def decompose(self, x: int, base_log: int, level: int) -> np.ndarray:
        out = np.zeros(level, np.int64)
        self.lib.orc_decompose(C.c_uint64(x), C.c_int(base_log), C.c_int(level), _p(out, C.c_int64))
        return out

    def modswitch(self, x: int) -> int:
        lg = int(self.p.N * 2).bit_length() - 1
        return int(self.lib.orc_modswitch(C.c_uint64(x), C.c_int(lg)))

    def keyswitch(self, keys: Keys, cts: np.ndarray) -> np.ndarray:
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, self.big)
        out = np.zeros((cts.shape[0], self.small), np.uint64)
        self.lib.orc_keyswitch_batch(C.byref(self.p), _p(keys.ksk, C.c_uint64), _p(cts, C.c_uint64),
                                     C.c_int(cts.shape[0]), _p(out, C.c_uint64))
        return out

    def lut_poly(self, table) -> np.ndarray:
        t = np.asarray(table, np.uint8)
        assert len(t) == 1 << (63 - self.p.delta_log)
        out = np.zeros(self.p.N, np.uint64)
        self.lib.orc_lut_poly(C.c_int(self.p.N), _p(t, C.c_uint8), C.c_int(self.p.delta_log), _p(out, C.c_uint64))
        return out

    def lut_post(self, table) -> int:
        """the constant added to the body of a PBS result: delta/2 for a half-step table (entries 0x80 | e stand for
        e - 1/2; include/fhestr_engine.h, fhestr_lut_register), else 0"""
        return (1 << (self.p.delta_log - 1)) if (int(np.asarray(table)[0]) & 0x80) else 0

    def blind_rotate_exact(self, keys: Keys, ks_ct: np.ndarray, lut: np.ndarray) -> np.ndarray:
        acc = np.zeros((2, self.p.N), np.uint64)
        self.lib.orc_blind_rotate_exact(C.byref(self.p), _p(keys.bsk, C.c_uint64),
                                        _p(np.ascontiguousarray(ks_ct), C.c_uint64),
                                        _p(np.ascontiguousarray(lut), C.c_uint64), _p(acc, C.c_uint64))
        return acc

    def sample_extract(self, acc: np.ndarray) -> np.ndarray:
        out = np.zeros(self.big, np.uint64)
        self.lib.orc_sample_extract(C.c_int(self.p.N), _p(np.ascontiguousarray(acc), C.c_uint64), _p(out, C.c_uint64))
        return out

    def external_product_exact(self, ggsw: np.ndarray, glwe: np.ndarray, acc: np.ndarray) -> np.ndarray:
        acc = np.ascontiguousarray(acc, np.uint64).copy()
        self.lib.orc_external_product_exact(C.byref(self.p), _p(np.ascontiguousarray(ggsw), C.c_uint64),
                                            _p(np.ascontiguousarray(glwe), C.c_uint64), _p(acc, C.c_uint64))
        return acc

    def fourier_bsk(self, keys: Keys) -> np.ndarray:
        out = np.zeros(int(self.lib.orc_fourier_bsk_doubles(C.byref(self.p))), np.float64)
        self.lib.orc_fourier_bsk(C.byref(self.p), _p(keys.bsk, C.c_uint64), _p(out, C.c_double))
        return out

    def external_product_fft(self, ggsw_f: np.ndarray, glwe: np.ndarray, acc: np.ndarray) -> np.ndarray:
        acc = np.ascontiguousarray(acc, np.uint64).copy()
        self.lib.orc_external_product_fft(C.byref(self.p), _p(np.ascontiguousarray(ggsw_f), C.c_double),
                                          _p(np.ascontiguousarray(glwe), C.c_uint64), _p(acc, C.c_uint64))
        return acc

    @staticmethod
    def _add_post(out, ids, post):
        if post is not None:
            with np.errstate(over="ignore"):
                out[:, -1] += np.asarray(post, np.uint64)[np.asarray(ids, np.int64)]
        return out

    def pbs_exact(self, keys: Keys, luts: np.ndarray, lut_ids, cts: np.ndarray, post=None) -> np.ndarray:
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, self.big)
        luts = np.ascontiguousarray(luts, np.uint64).reshape(-1, self.p.N)
        ids = np.ascontiguousarray(lut_ids, np.int32)
        out = np.zeros_like(cts)
        self.lib.orc_pbs_exact_batch(C.byref(self.p), _p(keys.bsk, C.c_uint64), _p(keys.ksk, C.c_uint64),
                                     _p(luts, C.c_uint64), _p(ids, C.c_int32), _p(cts, C.c_uint64),
                                     C.c_int(cts.shape[0]), _p(out, C.c_uint64))
        return self._add_post(out, ids, post)

    def pbs_fft(self, keys: Keys, fbsk: np.ndarray, luts: np.ndarray, lut_ids, cts: np.ndarray, post=None):
        """f64-FFT PBS (tfhe-rs' own route), OpenMP across ciphertexts.  Returns (out, threads)."""
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, self.big)
        luts = np.ascontiguousarray(luts, np.uint64).reshape(-1, self.p.N)
        ids = np.ascontiguousarray(lut_ids, np.int32)
        out = np.zeros_like(cts)
        th = self.lib.orc_pbs_fft_batch(C.byref(self.p), _p(fbsk, C.c_double), _p(keys.ksk, C.c_uint64),
                                        _p(luts, C.c_uint64), _p(ids, C.c_int32), _p(cts, C.c_uint64),
                                        C.c_int(cts.shape[0]), _p(out, C.c_uint64))
        return self._add_post(out, ids, post), int(th)
