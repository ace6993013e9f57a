"""ctypes binding of the client half of the C ABI (fhestr_client_*): the stand-in for MyClientKey
(/root/reference/src/client_key.rs:9-106).  Host-side, like the reference's client."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import EngineError, Params, PARAM_MESSAGE_2_CARRY_2_KS_PBS, load_library

# noise of PARAM_MESSAGE_2_CARRY_2_KS_PBS
LWE_STD = 7.069849454709433e-6
GLWE_STD = 2.9403601535432533e-16


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class ClientKey:
    """MyClientKey::from_params (client_key.rs:30-35): secret keys + the server key material."""

    def __init__(self, seed: int | None = None, lwe_std: float = LWE_STD, glwe_std: float = GLWE_STD, **params):
        """seed=None (default): keys and all encryption randomness from OS entropy.  An explicit non-zero seed gives
        reproducible keys and is for tests and benchmarks only."""
        if seed is not None and int(seed) == 0:
            raise ValueError("seed 0 is reserved (it means OS entropy at the C ABI); pass None for that")
        self.lib = load_library()
        prm = dict(PARAM_MESSAGE_2_CARRY_2_KS_PBS)
        prm.update(params)
        self.params = Params(**prm)
        self.n, self.N = self.params.n, self.params.N
        self.big = self.N + 1
        h = C.c_void_p()
        rc = self.lib.fhestr_client_create(C.byref(self.params), C.c_double(lwe_std), C.c_double(glwe_std),
                                           C.c_uint64(0 if seed is None else int(seed)), C.byref(h))
        if rc:
            raise EngineError(f"fhestr_client_create failed ({rc})")
        self.h = h
        self.lib.fhestr_client_destroy.restype = None
        self.lib.fhestr_client_destroy.argtypes = [C.c_void_p]

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhestr_client_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise EngineError(f"client call failed ({rc})")

    def server_keys(self):
        """(bsk_std [n][1][2][2][N], ksk [N][ks_level][n+1]) -- what get_server_key() hands over"""
        p = self.params
        bsk = np.zeros((p.n, p.pbs_level, 2, 2, p.N), np.uint64)
        ksk = np.zeros((p.N, p.ks_level, p.n + 1), np.uint64)
        self._ck(self.lib.fhestr_client_server_keys(self.h, _u64p(bsk), _u64p(ksk)))
        return bsk, ksk

    def secret_keys(self):
        s_lwe, s_glwe = np.zeros(self.n, np.uint8), np.zeros(self.N, np.uint8)
        self._ck(self.lib.fhestr_client_secret_keys(self.h, _u8p(s_lwe), _u8p(s_glwe)))
        return s_lwe, s_glwe

    def encrypt_blocks(self, values) -> np.ndarray:
        v = np.ascontiguousarray(values, np.uint8).ravel()
        out = np.zeros((len(v), self.big), np.uint64)
        self._ck(self.lib.fhestr_client_encrypt_blocks(self.h, _u8p(v), C.c_uint32(len(v)), _u64p(out)))
        return out

    def decrypt_blocks(self, cts: np.ndarray, with_error: bool = False):
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, self.big)
        vals = np.zeros(cts.shape[0], np.uint8)
        err = np.zeros(cts.shape[0], np.int64) if with_error else None
        self._ck(self.lib.fhestr_client_decrypt_blocks(
            self.h, _u64p(cts), C.c_uint32(cts.shape[0]), _u8p(vals),
            err.ctypes.data_as(C.POINTER(C.c_int64)) if with_error else None))
        return (vals, err) if with_error else vals

    def encrypt_u8(self, data) -> np.ndarray:
        """bytes -> [count][4][N+1]: FheAsciiChar::encrypt (fheasciichar.rs:27)"""
        b = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        out = np.zeros((len(b), 4, self.big), np.uint64)
        self._ck(self.lib.fhestr_client_encrypt_u8(self.h, _u8p(b), C.c_uint32(len(b)), _u64p(out)))
        return out

    def decrypt_u8(self, cts: np.ndarray) -> np.ndarray:
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, 4, self.big)
        out = np.zeros(cts.shape[0], np.uint8)
        self._ck(self.lib.fhestr_client_decrypt_u8(self.h, _u64p(cts), C.c_uint32(cts.shape[0]), _u8p(out)))
        return out

    def encrypt_str(self, s: str, padding: int = 0) -> np.ndarray:
        """MyClientKey::encrypt (client_key.rs:45-65): ASCII, no NUL, `padding` NULs appended"""
        assert all(0 < ord(ch) < 128 for ch in s), "The input string must only contain ascii letters and not include null characters"
        return self.encrypt_u8((s + "\0" * padding).encode("ascii"))

    def decrypt_str(self, cts: np.ndarray) -> str:
        """MyClientKey::decrypt (client_key.rs:96-106): truncate at the first NUL"""
        b = bytes(self.decrypt_u8(cts))
        cut = b.find(b"\0")
        return (b if cut < 0 else b[:cut]).decode("utf-8")
