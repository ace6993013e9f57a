"""End-to-end parity on the GPU through the reference-style API (fhestring_b200/fhestring.py over the C ABI):
encrypt with the real PARAM_MESSAGE_2_CARRY_2_KS_PBS parameters, run the string method as batched PBS levels
on the B200, decrypt, and compare with (i) the plaintext oracle of the REFERENCE algorithm and (ii) what Rust
`std` returns (the reference's own unit-test expectation, tests/golden/reference_tests.json).  Integer
results: bit-exact."""
import random

import numpy as np
import pytest

from oracle import fhestring_plain as P
from strcases import SIGNATURES, decode_result, encode_args, reference_cases

pytestmark = pytest.mark.gpu

ORACLE_FN = {m: getattr(P, "length" if m == "len" else m) for m in SIGNATURES}


@pytest.fixture(scope="module")
def keys(build_lib):
    from fhestring_b200.fhestring import MyClientKey
    ck = MyClientKey.from_params(seed=20)
    sk = ck.get_server_key(arena_blocks=1 << 19)
    pp = ck.get_public_parameters()
    yield ck, sk, pp
    sk.engine.close()


def call_method(ck, sk, pp, method, args, padding):
    """the reference test body: encrypt the arguments, call the server-key method, decrypt"""
    kinds, rkind = SIGNATURES[method]
    enc = []
    for kind, a in zip(kinds, args):
        if kind == "s":
            enc.append(ck.encrypt(a, padding, pp, sk.key))
        elif kind == "p":
            enc.append(ck.encrypt_no_padding(a))
        elif kind == "n":
            enc.append(ck.encrypt_char(int(a)))
        else:
            enc.append(int(a))
    res = getattr(sk, method)(*enc, pp)
    if rkind == "split":   # FheSplit: every buffer padded, plus the pattern_found flag, in one flush
        chars = [c for b in res.buffers for c in b.bytes] + [res.pattern_found]
        vals = ck.decrypt_padded(chars)
        bufs, k = [], 0
        for b in res.buffers:
            bufs.append(vals[k:k + len(b)])
            k += len(b)
        return bufs, vals[-1]
    if rkind == "u8":
        return ck.decrypt_char(res)
    if rkind == "str":
        return ck.decrypt_padded(res)
    return ck.decrypt_padded(res.string), ck.decrypt_char(res.pattern_found)


@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_reference_unit_tests_fast(keys, case):
    ck, sk, pp = keys
    sk.reset()
    sk.fast = True
    m = case["method"]
    if isinstance(case["expect"], str) and case["expect"].startswith("panic"):
        with pytest.raises(RuntimeError, match="Maximum supported size for find reached"):
            call_method(ck, sk, pp, m, case["args"], case["padding"])
        return
    got = call_method(ck, sk, pp, m, case["args"], case["padding"])
    assert decode_result(m, got) == case["expect"]                      # Rust std
    ref = ORACLE_FN[m](*encode_args(m, case["args"], case["padding"]))  # the reference's algorithm
    if SIGNATURES[m][1] == "str":
        assert list(got) == list(ref)
    elif SIGNATURES[m][1] == "u8":
        assert int(got) == int(ref)
    elif SIGNATURES[m][1] == "split":
        assert [list(b) for b in got[0]] == [list(b) for b in ref[0]] and int(got[1]) == int(ref[1])


def test_split_decrypt_like_the_reference(keys):
    """FheSplit::decrypt + trim_vector exactly as the reference's test body does (main.rs:931-953)"""
    from fhestring_b200.fhestring import FheSplit
    ck, sk, pp = keys
    sk.reset()
    my_string = ck.encrypt(" Mary had a", 1, pp, sk.key)
    pattern = ck.encrypt_no_padding(" ")
    plain_split, found = FheSplit.decrypt(sk.split(my_string, pattern, pp), ck)
    while plain_split and plain_split[0] == "":
        plain_split.pop(0)
    while plain_split and plain_split[-1] == "":
        plain_split.pop()
    assert plain_split == [x for x in " Mary had a".split(" ") if x] and found == 1
    sk.reset()
    got, found = FheSplit.decrypt(sk.rsplitn_clear(ck.encrypt(".A.B.C.", 1, pp, sk.key), ".", 3, pp), ck)
    assert [x for x in got if x] == ["C", ".A.B"]


FAITHFUL = ["valid_contains", "valid_ends_with", "valid_starts_with", "lowercase", "len", "find", "eq", "less_than",
            "greater_equal", "cli_Replace", "cli_StripSuffix", "cli_TrimStart", "cli_Concatenate", "cli_Rfind"]


@pytest.mark.parametrize("case", [c for c in reference_cases() if c["name"] in FAITHFUL], ids=lambda c: c["name"])
def test_reference_unit_tests_faithful(keys, case):
    """the reference's own op order (hundreds of tiny levels): same decrypted result"""
    ck, sk, pp = keys
    sk.reset()
    sk.fast = False
    try:
        got = call_method(ck, sk, pp, case["method"], case["args"], case["padding"])
    finally:
        sk.fast = True
    assert decode_result(case["method"], got) == case["expect"]


def test_char_primitives_eager(keys):
    """FheAsciiChar methods called one at a time, as the reference's sequential callers do (fheasciichar.rs:35-104)"""
    from fhestring_b200.fhestring import FheAsciiChar
    ck, sk, pp = keys
    sk.reset()
    a, b = ck.encrypt_char(0x61), ck.encrypt_char(0x7A)
    one = FheAsciiChar.encrypt_trivial(1, pp, sk)
    assert ck.decrypt_char(a.eq(sk, b)) == 0
    assert ck.decrypt_char(a.ne(sk, b)) == 1
    assert ck.decrypt_char(a.lt(sk, b)) == 1
    assert ck.decrypt_char(a.ge(sk, b)) == 0
    s = a.add(sk, b)
    assert ck.decrypt_char(s) == (0x61 + 0x7A) & 255
    d = a.sub(sk, b)
    assert ck.decrypt_char(d) == (0x61 - 0x7A) & 255
    # results of earlier executions feed later ones (commit keeps them in the arena)
    assert ck.decrypt_char(s.sub(sk, d)) == ((0x61 + 0x7A) - (0x61 - 0x7A)) & 255
    assert ck.decrypt_char(a.lt(sk, b).if_then_else(sk, a, b)) == 0x61
    assert ck.decrypt_char(a.eq(sk, b).bitor(sk, one)) == 1
    assert ck.decrypt_char(a.is_lowercase(sk)) == 1 and ck.decrypt_char(a.is_uppercase(sk)) == 0
    assert ck.decrypt_char(ck.encrypt_char(0x0A).is_whitespace(sk)) == 1
    assert ck.decrypt_char(a.eq(sk, a).flip(sk)) == 0


def test_config3_eq_ge_le_64_chars(keys):
    """BASELINE config 3: == / >= / <= on two 64-char strings (padding 1)"""
    ck, sk, pp = keys
    rng = random.Random(3)
    a = "".join(chr(rng.randrange(0x20, 0x7F)) for _ in range(64))
    variants = {
        "equal": a,
        "differ_at_0": chr(0x20 + (ord(a[0]) - 0x1F) % 0x5F) + a[1:],
        "differ_at_63": a[:63] + chr(0x20 + (ord(a[63]) - 0x1F) % 0x5F),
        "prefix_48": a[:48],
    }
    for name, b in variants.items():
        for m in ("eq", "ge", "le"):
            sk.reset()
            got = call_method(ck, sk, pp, m, [a, b], 1)
            std = {"eq": a == b, "ge": a >= b, "le": a <= b}[m]
            ref = ORACLE_FN[m](*encode_args(m, [a, b], 1))
            assert got == int(std) == ref, (name, m)


def test_config4_contains_find_256_chars(keys):
    """BASELINE config 4: encrypted 8-char pattern over a 256-char encrypted string (L = 257 < 255 + 8)"""
    ck, sk, pp = keys
    rng = random.Random(4)
    pat = "qzjxkvwq"
    for where in (0, 124, 248, None):
        body = [rng.choice("abcdefghilmnoprstu") for _ in range(256)]
        if where is not None:
            body[where:where + 8] = pat
        s = "".join(body)
        assert (s.find(pat) if where is not None else -1) == (where if where is not None else -1)
        sk.reset()
        es, ep = ck.encrypt(s, 1, pp, sk.key), ck.encrypt_no_padding(pat)
        c, f = sk.contains(es, ep, pp), sk.find(es, ep, pp)
        assert ck.decrypt_char(c) == int(where is not None)
        assert ck.decrypt_char(f) == (where if where is not None else 255)
        assert sk.last_info.n_levels <= 10


def test_plan_cache_reruns_a_query_shape_on_new_inputs(keys):
    """the second call of a string method on fresh ciphertexts of the same lengths (right after reset()) re-runs the bound
    program of the first instead of recording again: same results as Rust std on every call, results of a cached run
    usable as operands of further ops, and plan_cache = False records every time"""
    ck, sk, pp = keys
    rng = random.Random(11)
    hits0 = sk.plan_hits
    words = ["".join(rng.choice("abcdefgh") for _ in range(12)) for _ in range(4)]
    pats = ["cab", words[1][4:7], "hhh", words[3][9:12]]
    for w, p in zip(words, pats):                       # contains + a follow-up op on the same inputs
        sk.reset()
        es, ep = ck.encrypt(w, 2, pp, sk.key), ck.encrypt_no_padding(p)
        c = sk.contains(es, ep, pp)
        assert ck.decrypt_char(c) == int(p in w), (w, p)
        assert ck.decrypt_char(sk.starts_with(es, ep, pp)) == int(w.startswith(p)), (w, p)
    assert hits0 + 3 <= sk.plan_hits <= hits0 + 4       # the first call of a shape records, the others re-run its program
    for w in words:                                     # a string result of a cached run as the operand of another method
        sk.reset()
        up = sk.to_upper(ck.encrypt(w, 2, pp, sk.key), pp)
        low = sk.to_lower(up, pp)
        assert ck.decrypt(up) == w.upper() and ck.decrypt(low) == w
    assert hits0 + 6 <= sk.plan_hits <= hits0 + 8
    for a, b in (("hello", "hello"), ("hellp", "hello"), ("abcde", "abcdf")):
        sk.reset()
        ea, eb = ck.encrypt(a, 1, pp, sk.key), ck.encrypt(b, 1, pp, sk.key)
        assert ck.decrypt_char(sk.eq(ea, eb, pp)) == int(a == b)
        sk.reset()
        ea, eb = ck.encrypt(a, 1, pp, sk.key), ck.encrypt(b, 1, pp, sk.key)
        assert ck.decrypt_char(sk.ge(ea, eb, pp)) == int(a >= b)
    sk.plan_cache = False
    try:
        h = sk.plan_hits
        sk.reset()
        assert ck.decrypt_char(sk.eq(ck.encrypt("hello", 1, pp, sk.key), ck.encrypt("hello", 1, pp, sk.key), pp)) == 1
        assert sk.plan_hits == h
    finally:
        sk.plan_cache = True


def test_config5_replace_1024_chars(keys):
    """BASELINE config 5: replace with encrypted from/to (4 chars each) over a 1024-char padded string"""
    ck, sk, pp = keys
    rng = random.Random(5)
    body = [rng.choice("abcdfghijkmnpqrstuvwxyz ") for _ in range(1000)]
    for pos in (0, 100, 333, 500, 640, 801, 900, 996):
        body[pos:pos + 4] = "ello"
    s = "".join(body)
    sk.reset()
    es = ck.encrypt(s, 24, pp, sk.key)
    res = sk.replace(es, ck.encrypt_no_padding("ello"), ck.encrypt_no_padding("_llo"), pp)
    got = ck.decrypt_padded(res)
    assert len(got) == 1025
    assert P.decrypt_str(got) == s.replace("ello", "_llo")
    assert got == P.replace(encode_args("replace", [s, "ello", "_llo"], 24)[0], [ord(c) for c in "ello"], [ord(c) for c in "_llo"])


def test_noise_of_deep_graph_outputs(keys):
    """phase error of the result blocks of a long dependent chain stays inside the decoding margin (delta/2 = 2^58)
    with room to spare: |err| < 2^55 (an 8-sigma bound for the parameter set's PBS output noise 2^-15.5 is 2^51.5)"""
    ck, sk, pp = keys
    sk.reset()
    s = ck.encrypt("hello world", 1, pp, sk.key)
    up = sk.to_upper(sk.to_lower(sk.to_upper(s, pp), pp), pp)
    cts = ck._host_cts(up.bytes)
    vals, err = ck.client.decrypt_blocks(cts.reshape(-1, ck.client.big), with_error=True)
    assert ck.decrypt(up) == "HELLO WORLD"
    assert np.abs(err).max() < (1 << 55)


def test_every_arena_block_matches_the_plaintext_interpretation(keys):
    """not only the decrypted outputs: EVERY block the compiled program writes (all intermediate PBS results and
    leveled sums, padding-bit values included mod 16) equals the plaintext interpretation of the same job list"""
    from fhestring_b200.graph import Graph
    from plain_exec import blocks_of, run_program
    ck, sk, pp = keys
    sk.reset()
    eng = sk.engine
    for method, args in (("find", ["the quick brown fox", "brown"]), ("replace", ["hello world world", "world", "abc"]),
                         ("ge", ["straw", "strap"]), ("trim", [" \tpadded \n"])):
        enc = encode_args(method, args, 1)
        g = Graph()
        ids, slots, vals = [], [], []
        for a in enc:
            i, s = g.input_chars(len(a))
            ids.append(i); slots.append(s.reshape(-1)); vals.append(blocks_of(a).reshape(-1))
        rs, rc = g.string_op(method, ids, fast=True)
        outs = ([] if rs is None else list(rs)) + ([] if rc is None else [rc])
        g.mark_output(outs)
        info = g.compile(1)
        in_slots, in_vals = np.concatenate(slots), np.concatenate(vals)
        eng.upload(int(in_slots[0]), ck.client.encrypt_blocks(in_vals.astype(np.uint8)))
        plain = run_program(g, in_slots, in_vals)
        jobs, off, npbs, first = g.program()
        g.execute(eng)
        got = ck.client.decrypt_blocks(eng.download(0, info.slots_used)).astype(np.int64)
        written = np.array([int(j["dst"]) for j in jobs], np.int64)
        assert np.array_equal(got[written], plain[written] % 16), method
        g.close()


@pytest.mark.parametrize("fast", [True, False], ids=["fast", "reference_op_order"])
def test_config1_cli_all_52_methods(keys, fast):
    """BASELINE config 1: `fhestring --string hello --pattern ello --n 1 --from ello --to _llo`, every algorithm the
    CLI runs (/root/reference/src/main.rs:47-115: 52 methods, the 18 *_clear forms included, in the CLI's order), each
    result bit-exact against Rust std AND against the plaintext restatement of the reference's algorithm"""
    from cli_config1 import METHODS, run_all
    ck, sk, pp = keys
    sk.fast = fast
    try:
        rows = run_all(ck, sk, pp, "hello", "ello", 1, "ello", "_llo")
    finally:
        sk.fast = True
    assert [r["method"] for r in rows] == METHODS
    bad = [(r["method"], r["got"], r["std"], r["oracle"]) for r in rows if not r["passed"]]
    assert not bad, bad


def test_config1_cli_second_input_set(keys):
    """the same loop on inputs where the pattern occurs twice, n = 2 and from/to differ in length"""
    from cli_config1 import run_all
    ck, sk, pp = keys
    rows = run_all(ck, sk, pp, "a bc bc d", "bc", 2, "bc", "xyz")
    bad = [(r["method"], r["got"], r["std"], r["oracle"]) for r in rows if not r["passed"]]
    assert not bad, bad
