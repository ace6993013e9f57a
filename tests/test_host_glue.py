"""Host-side glue of the reference-facing mirror (fhestring_b200/fhestring.py) that needs no GPU: a string's ciphertexts that
lie back to back in memory are uploaded as they lie (no gathering copy), anything else is gathered."""
import types

import numpy as np

from fhestring_b200.fhestring import FheAsciiChar, MyServerKey


def _sk(big=2049):
    sk = MyServerKey.__new__(MyServerKey)          # no engine: only the layout logic is under test
    sk.engine = types.SimpleNamespace(big=big)
    return sk


def test_contiguous_views_are_uploaded_in_place():
    big = 2049
    base = np.arange(7 * 4 * big, dtype=np.uint64).reshape(7, 4, big)
    sk = _sk(big)
    chars = [FheAsciiChar(ct=base[i]) for i in range(2, 6)]
    run = sk._contiguous_run(chars)
    assert run is not None and run.shape == (16, big)
    assert np.shares_memory(run, base) and np.array_equal(run, base[2:6].reshape(-1, big))
    flat = base.reshape(-1, big)                      # views of a 2-D view of the same buffer
    chars = [FheAsciiChar(ct=flat[4 * i:4 * i + 4]) for i in range(3)]
    run = sk._contiguous_run(chars)
    assert run is not None and np.array_equal(run, flat[:12])


def test_everything_else_is_gathered():
    big = 2049
    base = np.zeros((6, 4, big), np.uint64)
    sk = _sk(big)
    assert sk._contiguous_run([FheAsciiChar(ct=base[0]), FheAsciiChar(ct=base[2])]) is None            # a gap
    assert sk._contiguous_run([FheAsciiChar(ct=base[1]), FheAsciiChar(ct=base[0])]) is None            # out of order
    assert sk._contiguous_run([FheAsciiChar(ct=base[0].astype(np.int64))]) is None                      # wrong dtype
    assert sk._contiguous_run([FheAsciiChar(ct=base[:, :, ::2][0])]) is None                            # not contiguous / wrong shape
