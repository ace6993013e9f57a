"""Pin the plaintext oracle (oracle/fhestring_plain.py) against the reference's own unit tests and CLI
self-check: same literal inputs, expected value = what Rust `std` returns (tests/golden/reference_tests.json)."""
import pytest

from oracle import fhestring_plain as P
from strcases import SIGNATURES, decode_result, encode_args, reference_cases

ORACLE_FN = {m: getattr(P, "length" if m == "len" else m) for m in SIGNATURES}


@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_reference_unit_tests(case):
    m = case["method"]
    args = encode_args(m, case["args"], case["padding"])
    if isinstance(case["expect"], str) and case["expect"].startswith("panic"):
        with pytest.raises(P.FindTooLong):
            ORACLE_FN[m](*args)
        return
    assert decode_result(m, ORACLE_FN[m](*args)) == case["expect"]


def test_known_divergences_from_std():
    # the oracle models the reference, not std
    s = [ord("a")] * 3 + [0]
    assert P.decrypt_str(P.replace(s, [97, 97], [98])) == "bb"          # overlapping matches all fire
    assert P.length([65] * 256 + [0]) == 0                               # u8 wrapping
    assert P.find([97, 98, 0], [122]) == 255                             # not found
    assert P.bubble_zeroes_right([0, 1, 0, 2, 0, 3]) == [1, 2, 3, 0, 0, 0]
