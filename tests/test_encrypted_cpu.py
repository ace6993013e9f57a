"""Graph recordings executed on REAL ciphertexts at the real parameter set, on the CPU: the oracle's keyswitch and
f64-FFT blind rotation stand in for the GPU kernels (tests/tools/cpu_encrypted_exec.py).  The plaintext interpretation
checks values; this checks that the noise the recordings accumulate stays far inside the decoding margin."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
from cpu_encrypted_exec import execute  # noqa: E402
from plain_exec import blocks_of  # noqa: E402


@pytest.fixture(scope="module")
def Graph(build_lib):
    from fhestring_b200.graph import Graph
    return Graph


@pytest.fixture(scope="module")
def real_keys():
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    o = Oracle(**P)
    keys = o.keygen(1)
    return o, keys, o.fourier_bsk(keys)


def _run(Graph, real_keys, build):
    o, keys, fbsk = real_keys
    g = Graph()
    ids, slots, vals = [], [], []

    def add(chars):
        i, s = g.input_chars(len(chars))
        ids.append(i); slots.append(s.reshape(-1)); vals.append(blocks_of(list(chars)).reshape(-1))

    outs = build(g, add, ids)
    g.mark_output(outs)
    g.compile(1)
    got, plain, worst = execute(o, keys, fbsk, g, np.concatenate(slots), np.concatenate(vals))
    jobs, off, npbs, first = g.program()
    bad = [int(j["dst"]) for j in jobs if got[int(j["dst"])] != plain[int(j["dst"])] % 16]
    g.close()
    return bad, worst, len(jobs)


def test_shallow_comparison_on_real_ciphertexts(Graph, real_keys):
    rng = np.random.default_rng(8)
    a = [int(x) for x in rng.integers(32, 127, 32)] + [0]
    b = list(a)
    b[20] ^= 1

    def build(g, add, ids):
        add(a); add(b)
        return [g.string_op("ge", ids, fast=True)[1]]

    bad, worst, n = _run(Graph, real_keys, build)
    assert not bad and n > 100
    assert worst < 0.1, worst   # in units of delta; the decoding margin is 0.5, keyswitch + mod-switch add sigma 0.07


def test_depth_minimised_split_on_real_ciphertexts(Graph, real_keys):
    def build(g, add, ids):
        add([ord(c) for c in "hello"] + [0]); add([ord(c) for c in "ello"])
        bufs, found = g.split_op("split", ids, fast=True)
        return [x for bb in bufs for x in bb] + [found]

    bad, worst, n = _run(Graph, real_keys, build)
    assert not bad and n > 300
    assert worst < 0.1, worst
