"""Reader / writer of the flat "FHESTRFX" v1 file that `integration/fhestr-parity` (Rust, tfhe-rs 0.5.2) produces:
the key hand-over in the layouts `fhestr_load_keys` takes, and golden (input LWE, keyswitched LWE, PBS output) triples
plus LUT accumulators out of tfhe-rs itself -- the vectors that pin keyswitch / LUT generation / sample extract at the
ciphertext level (tests/test_tfhe_rs_fixture.py).  Layout: integration/fhestr-parity/src/main.rs, module docstring.

Replaces holding a `tfhe::integer::ServerKey` (/root/reference/src/server_key/mod.rs:13-16, created at
/root/reference/src/client_key.rs:31-39): `load(path)` -> `Engine.load_keys(fx.bsk_std, fx.ksk)`.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC = b"FHESTRFX"
PARAM_FIELDS = ("n", "N", "k", "pbs_base_log", "pbs_level", "ks_base_log", "ks_level", "delta_log")


@dataclass
class Fixture:
    params: dict
    bsk_std: np.ndarray                      # [n][pbs_level][k+1][k+1][N] u64
    ksk: np.ndarray                          # [k*N][ks_level][n+1] u64
    s_lwe: np.ndarray | None = None          # [n] u8
    s_glwe: np.ndarray | None = None         # [k*N] u8
    tables: np.ndarray = field(default_factory=lambda: np.zeros((0, 16), np.uint8))
    lut_bodies: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.uint64))
    triple_lut: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    triple_in: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.uint64))
    triple_ks: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.uint64))
    triple_out: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), np.uint64))


def load(path: str) -> Fixture:
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise ValueError(f"{path}: not a FHESTRFX file")
    version, = struct.unpack_from("<I", raw, 8)
    if version != 1:
        raise ValueError(f"{path}: FHESTRFX version {version}, this reader knows 1")
    vals = struct.unpack_from("<8i3I", raw, 12)
    p = dict(zip(PARAM_FIELDS, vals[:8]))
    has_secrets, n_luts, n_triples = vals[8:]
    off = 12 + 8 * 4 + 3 * 4
    n, N, k = p["n"], p["N"], p["k"]
    big = k * N

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype, count, off).copy()
        off += a.nbytes
        return a

    fx = Fixture(params=p, bsk_std=None, ksk=None)
    if has_secrets:
        fx.s_lwe, fx.s_glwe = take(np.uint8, n), take(np.uint8, big)
    fx.bsk_std = take("<u8", n * p["pbs_level"] * (k + 1) * (k + 1) * N).reshape(n, p["pbs_level"], k + 1, k + 1, N)
    fx.ksk = take("<u8", big * p["ks_level"] * (n + 1)).reshape(big, p["ks_level"], n + 1)
    tables, bodies = [], []
    for _ in range(n_luts):
        tables.append(take(np.uint8, 16))
        bodies.append(take("<u8", N))
    fx.tables = np.stack(tables) if tables else np.zeros((0, 16), np.uint8)
    fx.lut_bodies = np.stack(bodies) if bodies else np.zeros((0, N), np.uint64)
    tl, ti, tk, to = [], [], [], []
    for _ in range(n_triples):
        tl.append(take("<u4", 1)[0])
        ti.append(take("<u8", big + 1)); tk.append(take("<u8", n + 1)); to.append(take("<u8", big + 1))
    if n_triples:
        fx.triple_lut, fx.triple_in, fx.triple_ks, fx.triple_out = np.array(tl, np.uint32), np.stack(ti), np.stack(tk), np.stack(to)
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes (truncated or mis-sized file)")
    return fx


def save(path: str, fx: Fixture) -> None:
    """the same layout the Rust harness writes (used by the tests to build a synthetic file out of the oracle)"""
    p = fx.params
    has_secrets = fx.s_lwe is not None
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", 1) + struct.pack("<8i", *[int(p[k]) for k in PARAM_FIELDS]))
        f.write(struct.pack("<3I", int(has_secrets), len(fx.tables), len(fx.triple_lut)))
        if has_secrets:
            f.write(np.ascontiguousarray(fx.s_lwe, np.uint8).tobytes() + np.ascontiguousarray(fx.s_glwe, np.uint8).tobytes())
        f.write(np.ascontiguousarray(fx.bsk_std, "<u8").tobytes())
        f.write(np.ascontiguousarray(fx.ksk, "<u8").tobytes())
        for t, b in zip(fx.tables, fx.lut_bodies):
            f.write(np.ascontiguousarray(t, np.uint8).tobytes() + np.ascontiguousarray(b, "<u8").tobytes())
        for i in range(len(fx.triple_lut)):
            f.write(struct.pack("<I", int(fx.triple_lut[i])))
            for a in (fx.triple_in[i], fx.triple_ks[i], fx.triple_out[i]):
                f.write(np.ascontiguousarray(a, "<u8").tobytes())
