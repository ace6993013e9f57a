"""Best-effort reader of a bincode-serialised `tfhe::integer::ServerKey` (tfhe-rs 0.5.2), the object the reference
holds (/root/reference/src/server_key/mod.rs:13-16) and can serialise through the serde derive of
/root/reference/src/client_key.rs:9 -- producing the (bsk_std, ksk) pair `fhestr_load_keys` takes.

STATUS: UNVERIFIED AGAINST A REAL FILE.  There is no Rust toolchain in the build image, so no tfhe-rs-written bincode
exists here; the field order below is tfhe-rs 0.5.2's `SerializableServerKey` / `LweKeyswitchKey` /
`FourierLweBootstrapKey` / `FourierPolynomialList` as recalled, with bincode 1.x default options (little endian,
u64 lengths and usize, u32 enum tags).  Every structural assumption is CHECKED while parsing (lengths must multiply
out, tags must be known, the file must be consumed exactly), so a layout mismatch fails loudly instead of loading a
wrong key.  The SUPPORTED route is `integration/fhestr-parity export-keys`, where tfhe-rs itself deserialises the key
and inverse-transforms the Fourier bootstrapping key with its own FFT plan; this module exists so that a Python-only
deployment has a starting point, and tests/test_tfhe_rs_import.py pins it against a writer of the same layout.

Fourier bootstrapping key: tfhe-rs serialises every Fourier polynomial in concrete-fft's STANDARD order
(`serialize_fourier_buffer`), taken here as X_k = sum_j (p_j + i p_{j+N/2}) exp(i pi j / N) exp(-2 pi i j k / (N/2)) on
the coefficients read as signed integers.  The standard-domain key is the rounded inverse.
"""
from __future__ import annotations

import struct

import numpy as np


class BincodeError(ValueError):
    pass


class _Reader:
    def __init__(self, raw: bytes):
        self.raw, self.off = raw, 0

    def u32(self) -> int:
        v, = struct.unpack_from("<I", self.raw, self.off); self.off += 4; return v

    def u64(self) -> int:
        v, = struct.unpack_from("<Q", self.raw, self.off); self.off += 8; return v

    def u128(self) -> int:
        lo, hi = struct.unpack_from("<QQ", self.raw, self.off); self.off += 16; return lo | (hi << 64)

    def array(self, dtype, count: int) -> np.ndarray:
        a = np.frombuffer(self.raw, dtype, count, self.off)
        self.off += a.nbytes
        return a


def _ciphertext_modulus(r: _Reader):
    modulus, scalar_bits = r.u128(), r.u64()
    if scalar_bits != 64 or modulus not in (0, 1 << 64):
        raise BincodeError(f"unsupported ciphertext modulus {modulus} on {scalar_bits} bits (native 2^64 expected)")


def fourier_to_standard(spectrum: np.ndarray, N: int) -> np.ndarray:
    """[..., N/2] complex (standard order) -> [..., N] u64 torus words"""
    M = N // 2
    c = np.fft.ifft(spectrum, axis=-1) * np.exp(-1j * np.pi * np.arange(M) / N)
    p = np.concatenate([c.real, c.imag], axis=-1)
    return np.rint(p).astype(np.int64).astype(np.uint64)


def standard_to_fourier(poly: np.ndarray) -> np.ndarray:
    N = poly.shape[-1]
    M = N // 2
    s = poly.astype(np.int64).astype(np.float64)
    c = (s[..., :M] + 1j * s[..., M:]) * np.exp(1j * np.pi * np.arange(M) / N)
    return np.fft.fft(c, axis=-1)


def read_server_key(raw: bytes):
    """-> (params dict, bsk_std [n][level][k+1][k+1][N] u64, ksk [k*N][ks_level][n+1] u64)"""
    r = _Reader(raw)
    # ---- LweKeyswitchKey { data: Vec<u64>, decomp_base_log, decomp_level_count, output_lwe_size, ciphertext_modulus }
    ks_words = r.u64()
    ksk_flat = r.array("<u8", ks_words)
    ks_base_log, ks_level, out_size = r.u64(), r.u64(), r.u64()
    _ciphertext_modulus(r)
    n = out_size - 1
    if ks_level == 0 or ks_words % (ks_level * out_size):
        raise BincodeError("keyswitch key size does not factor as input_dim * level * output_lwe_size")
    big = ks_words // (ks_level * out_size)
    # ---- ShortintBootstrappingKey: enum tag 0 = Classic(FourierLweBootstrapKey)
    tag = r.u32()
    if tag != 0:
        raise BincodeError(f"bootstrapping key variant {tag}: only Classic (0) is supported (MultiBit keys are not)")
    # FourierPolynomialList: seq(2 + chunks) [polynomial_size, chunks, chunks x seq(N/2) of c64]
    seq_len = r.u64()
    N, chunks = r.u64(), r.u64()
    if seq_len != 2 + chunks or N < 2 or N & (N - 1):
        raise BincodeError("Fourier polynomial list header is inconsistent")
    spec = np.zeros((chunks, N // 2), np.complex128)
    for q in range(chunks):
        if r.u64() != N // 2:
            raise BincodeError("Fourier polynomial of unexpected length")
        spec[q] = r.array("<f8", N).view(np.complex128)
    in_dim, glwe_size, pbs_base_log, pbs_level = r.u64(), r.u64(), r.u64(), r.u64()
    k = glwe_size - 1
    if in_dim != n or chunks != n * pbs_level * glwe_size * glwe_size or big != k * N:
        raise BincodeError("bootstrapping key dimensions do not match the keyswitching key")
    message_modulus, carry_modulus, max_degree, max_noise_level = r.u64(), r.u64(), r.u64(), r.u64()
    _ciphertext_modulus(r)
    pbs_order = r.u32()
    if r.off != len(raw):
        raise BincodeError(f"{len(raw) - r.off} unread bytes: the recalled layout does not match this file")
    if pbs_order != 0:
        raise BincodeError("PBSOrder::BootstrapKeyswitch keys are not supported (the reference uses KeyswitchBootstrap)")
    # tfhe-rs keeps the decomposition levels of a GGSW in REVERSE order in memory (last level first)
    bsk = fourier_to_standard(spec, N).reshape(n, pbs_level, glwe_size, glwe_size, N)[:, ::-1]
    total_bits = (message_modulus * carry_modulus).bit_length() - 1
    params = dict(n=int(n), N=int(N), k=int(k), pbs_base_log=int(pbs_base_log), pbs_level=int(pbs_level),
                  ks_base_log=int(ks_base_log), ks_level=int(ks_level), delta_log=63 - total_bits)
    return params, np.ascontiguousarray(bsk), ksk_flat.reshape(big, ks_level, out_size).copy()


def write_server_key(params: dict, bsk_std: np.ndarray, ksk: np.ndarray, message_modulus=4, carry_modulus=4) -> bytes:
    """the inverse of read_server_key in the same recalled layout (tests; NOT a tfhe-rs-compatible writer by proof)"""
    n, N, k = params["n"], params["N"], params["k"]
    out = [struct.pack("<Q", ksk.size), np.ascontiguousarray(ksk, "<u8").tobytes(),
           struct.pack("<3Q", params["ks_base_log"], params["ks_level"], n + 1), struct.pack("<QQQ", 0, 0, 64),
           struct.pack("<I", 0)]
    spec = standard_to_fourier(np.ascontiguousarray(bsk_std)[:, ::-1].reshape(-1, N))
    out.append(struct.pack("<3Q", 2 + len(spec), N, len(spec)))
    for row in spec:
        out.append(struct.pack("<Q", N // 2) + np.ascontiguousarray(row).view(np.float64).astype("<f8").tobytes())
    out.append(struct.pack("<4Q", n, k + 1, params["pbs_base_log"], params["pbs_level"]))
    out.append(struct.pack("<4Q", message_modulus, carry_modulus, message_modulus * carry_modulus - 1,
                           (message_modulus * carry_modulus - 1) // (message_modulus - 1)))
    out.append(struct.pack("<QQQ", 0, 0, 64) + struct.pack("<I", 0))
    return b"".join(out)
