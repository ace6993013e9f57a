"""Interpret a compiled job list on plaintext block values -- what the engine does on ciphertexts, minus the
noise.  Values live mod 32 (bit 4 = padding bit).  A PBS on v >= 16 returns -f(v - 16);
v == 16 is ambiguous under noise, so reaching it is reported as an error."""
from __future__ import annotations

import numpy as np


def shard_range(n_jobs: int, rank: int, world: int):
    """restates fhestr_shard_range (include/fhestr_engine.h)"""
    per = (n_jobs + world - 1) // world
    lo = min(per * rank, n_jobs)
    hi = min(lo + per, n_jobs)
    return lo, hi, per


def run_jobs(values: np.ndarray, jobs, luts, delta_log: int = 59):
    """execute jobs in order on `values` (int array over arena slots, mod 32)"""
    for j in jobs:
        nt = int(j["n_terms"])
        v = int(j["constant"]) >> delta_log
        for t in range(nt):
            v += int(j["coeff"][t]) * int(values[int(j["src"][t])])
        v %= 32
        lut = int(j["lut"])
        if lut < 0:
            values[int(j["dst"])] = v
        else:
            e = int(luts[lut][v & 15])
            if e & 0x80:       # half-step table: f(v) = e[v] below 16, 1 - e[v - 16] from 16 on (fhestr_lut_register)
                e &= 0x7f
                values[int(j["dst"])] = e if v < 16 else (1 - e) % 32
            else:
                assert v != 16, "PBS input reached the ambiguous value 16"
                values[int(j["dst"])] = e if v < 16 else (-e) % 32


def run_program(graph, input_slots, input_blocks, values=None):
    """run the last compile of `graph` (fhestring_b200.graph.Graph) on plaintext.
    input_slots [k], input_blocks [k]: block values of the encrypted inputs.  Returns the slot value array."""
    info = graph.info
    jobs, off, npbs, first = graph.program()
    luts = graph.luts()
    if values is None:
        # unwritten arena slots hold garbage (whatever an earlier query left there), never a convenient 0
        values = np.random.default_rng(info.slots_used).integers(0, 32, info.slots_used).astype(np.int64)
    elif len(values) < info.slots_used:
        values = np.concatenate([values, np.zeros(info.slots_used - len(values), np.int64)])
    values[np.asarray(input_slots, np.int64)] = np.asarray(input_blocks, np.int64)
    tslots, tvals = graph.trivials()
    values[tslots.astype(np.int64)] = tvals
    for l in range(info.n_levels):
        a, b = int(off[l]), int(off[l + 1])
        # levels are sets of INDEPENDENT jobs: PBS jobs may not read what the same level writes
        written = set(int(j["dst"]) for j in jobs[a:a + int(npbs[l])])
        for j in jobs[a:a + int(npbs[l])]:
            for t in range(int(j["n_terms"])):
                assert int(j["src"][t]) not in written, "PBS job reads a result of its own level"
        run_jobs(values, jobs[a:b], luts)
    return values


def blocks_of(chars):
    """u8 values -> little-endian base-4 digits [len][4]"""
    c = np.asarray(chars, np.int64).reshape(-1, 1)
    return (c >> (2 * np.arange(4))) & 3


def chars_of(values, slots):
    """slot values [.., 4] -> u8 (blocks must be clean)"""
    b = values[np.asarray(slots, np.int64)]
    assert ((b >= 0) & (b <= 3)).all(), f"result blocks are not clean: {b}"
    return (b * (1 << (2 * np.arange(4)))).sum(-1)
