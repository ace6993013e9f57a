"""The op graph (fhestring_b200/csrc/graph.cpp, strings.cpp) against the plaintext oracle
(oracle/fhestring_plain.py): every char primitive and every string method, in both recordings (the
reference's op order and the depth-minimised one), compiled to PBS job levels and interpreted on plaintext
block values.  Host-only: no GPU."""
import random

import numpy as np
import pytest

from oracle import fhestring_plain as P
from plain_exec import blocks_of, chars_of, run_program
from strcases import SIGNATURES, decode_result, encode_args, reference_cases

ORACLE_FN = {m: getattr(P, "length" if m == "len" else m) for m in SIGNATURES}


@pytest.fixture(scope="module")
def Graph(build_lib):
    from fhestring_b200.graph import Graph
    return Graph


def run_method(Graph, method, enc_args, fast, slot_align=1):
    """record `method` on encrypted inputs holding enc_args, compile, interpret -> (decoded result, info)"""
    kinds, rkind = SIGNATURES[method]
    if rkind == "split":
        return run_split(Graph, method, enc_args, fast, slot_align)
    g = Graph()
    ids, in_slots, in_vals, clear_n = [], [], [], 0
    for kind, a in zip(kinds, enc_args):
        if kind == "c":
            clear_n = a
            continue
        vals = [a] if kind == "n" else list(a)
        i, s = g.input_chars(len(vals))
        ids.append(i)
        in_slots.append(s.reshape(-1))
        in_vals.append(blocks_of(vals).reshape(-1))
    rs, rc = g.string_op(method, ids, fast=fast, clear_n=clear_n)
    outs = ([] if rs is None else list(rs)) + ([] if rc is None else [rc])
    g.mark_output(outs)
    info = g.compile(slot_align)
    values = run_program(g, np.concatenate(in_slots) if in_slots else [], np.concatenate(in_vals) if in_vals else [])
    res_chars = chars_of(values, g.char_slots(outs)) if outs else np.zeros(0, np.int64)
    if rkind == "u8":
        return int(res_chars[0]), info
    if rkind == "str":
        return [int(v) for v in res_chars], info
    return ([int(v) for v in res_chars[:-1]], int(res_chars[-1])), info


def run_split(Graph, method, enc_args, fast, slot_align=1):
    kinds, _ = SIGNATURES[method]
    g = Graph()
    ids, in_slots, in_vals = [], [], []
    for kind, a in zip(kinds, enc_args):
        vals = [a] if kind == "n" else list(a)
        if not vals:           # empty pattern
            ids.append(np.zeros(0, np.uint32))
            continue
        i, s = g.input_chars(len(vals))
        ids.append(i)
        in_slots.append(s.reshape(-1))
        in_vals.append(blocks_of(vals).reshape(-1))
    bufs, found = g.split_op(method, ids, fast=fast)
    outs = list(bufs.reshape(-1)) + [found]
    g.mark_output(outs)
    info = g.compile(slot_align)
    values = run_program(g, np.concatenate(in_slots) if in_slots else [], np.concatenate(in_vals) if in_vals else [])
    res = chars_of(values, g.char_slots(outs))
    nb, bl = bufs.shape
    return ([[int(v) for v in res[b * bl:(b + 1) * bl]] for b in range(nb)], int(res[-1])), info


def oracle_raw(method, enc_args):
    return ORACLE_FN[method](*enc_args)


@pytest.mark.parametrize("fast", [0, 1], ids=["faithful", "fast"])
@pytest.mark.parametrize("case", reference_cases(), ids=lambda c: c["name"])
def test_reference_cases(Graph, case, fast):
    m = case["method"]
    args = encode_args(m, case["args"], case["padding"])
    if isinstance(case["expect"], str) and case["expect"].startswith("panic"):
        from fhestring_b200.engine import EngineError
        with pytest.raises(EngineError, match="Maximum supported size for find reached"):
            run_method(Graph, m, args, fast)
        return
    if not fast and m in ("repeat",) and len(args[0]) > 4:
        pytest.skip("16 x len chars through the O(L^2) bubble pass: covered on a shorter string below")
    if not fast and SIGNATURES[m][1] == "split" and len(args[0]) > 7:
        pytest.skip("L buffers through replace + the O(L^2) bubble pass each: the faithful recording is covered on the CLI inputs")
    got, _ = run_method(Graph, m, args, fast)
    assert decode_result(m, got) == case["expect"]
    # and the full padded result (not only the part before the first NUL) equals the reference algorithm's
    ref = oracle_raw(m, args)
    if SIGNATURES[m][1] == "str":
        assert list(got) == list(ref)
    elif SIGNATURES[m][1] == "strip":
        assert int(got[1]) == int(ref[1])
        if int(ref[1]) or m == "strip_prefix":
            assert list(got[0]) == list(ref[0])
    elif SIGNATURES[m][1] == "split":
        # every buffer, NULs included, and the pattern_found flag equal the reference algorithm's
        assert [list(b) for b in got[0]] == [list(b) for b in ref[0]]
        assert int(got[1]) == int(ref[1])


CHAR_CASES = [
    ("eq", 2, P.c_eq), ("ne", 2, P.c_ne), ("le", 2, P.c_le), ("lt", 2, P.c_lt), ("ge", 2, P.c_ge), ("gt", 2, P.c_gt),
    ("bitand", 2, P.c_and), ("bitor", 2, P.c_or), ("sub", 2, P.c_sub), ("add", 2, P.c_add),
    ("if_then_else", 3, P.c_ite), ("is_whitespace", 1, P.c_is_whitespace), ("is_uppercase", 1, P.c_is_uppercase),
    ("is_lowercase", 1, P.c_is_lowercase), ("flip", 1, P.c_flip),
]


@pytest.mark.parametrize("op,arity,fn", CHAR_CASES, ids=[c[0] for c in CHAR_CASES])
def test_char_primitives(Graph, op, arity, fn):
    """every FheAsciiChar primitive (fheasciichar.rs:35-168) on 96 operand tuples: edge values + random"""
    rng = random.Random(hash(op) & 0xFFFF)
    edge = [0, 1, 2, 3, 4, 15, 16, 31, 32, 64, 65, 90, 91, 96, 97, 122, 123, 127, 128, 254, 255, 9, 13, 14]
    tuples = [tuple(rng.choice(edge) for _ in range(arity)) for _ in range(48)]
    tuples += [tuple(rng.randrange(256) for _ in range(arity)) for _ in range(40)]
    tuples += [(v,) * arity for v in (0, 1, 255, 77)]
    if op == "flip":
        tuples = [(0,), (1,)] * 4   # flip is only ever applied to 0/1 flags (trim.rs:45, mod.rs:1508)... and any u8:
        tuples += [(v,) for v in (2, 77, 255)]
    g = Graph()
    ids, slots = g.input_chars(len(tuples) * arity)
    outs = []
    for k in range(len(tuples)):
        a = [int(ids[k * arity + i]) for i in range(arity)]
        outs.append(g.char_op(op, *a))
    g.mark_output(outs)
    g.compile()
    flat = [v for t in tuples for v in t]
    values = run_program(g, slots.reshape(-1), blocks_of(flat).reshape(-1))
    got = chars_of(values, g.char_slots(outs))
    assert [int(v) for v in got] == [fn(*t) for t in tuples]


def test_trivial_operands_fold(Graph):
    """PBS on trivial ciphertexts are evaluated in clear (tfhe-rs' trivial short cut)"""
    g = Graph()
    a, b = g.trivial_chars([200, 100])
    outs = [g.char_op("add", a, b), g.char_op("lt", a, b), g.char_op("if_then_else", a, b, a), g.char_op("eq", a, a)]
    g.mark_output(outs)
    info = g.compile()
    assert info.n_pbs == 0
    values = run_program(g, [], [])
    assert [int(v) for v in chars_of(values, g.char_slots(outs))] == [44, 0, 100, 1]


def rand_str(rng, n, alphabet="abAB z\t"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def random_cases(seed, count):
    rng = random.Random(seed)
    out = []
    for _ in range(count):
        m = rng.choice(sorted(k for k, v in SIGNATURES.items() if v[1] != "split"))
        kinds = SIGNATURES[m][0]
        n = rng.randrange(0, 7)
        s = rand_str(rng, n, "ab" if m in ("replace", "replacen", "contains", "find", "rfind") else "abAB z\t")
        args = []
        for i, k in enumerate(kinds):
            if k == "s":
                args.append(s if i == 0 else (s if rng.random() < 0.3 else rand_str(rng, rng.randrange(0, 6))))
            elif k == "p":
                plen = rng.randrange(0, 4)
                if i == 2 and m in ("replace", "replacen"):
                    plen = rng.randrange(0, len(args[1]) + 1) if rng.random() < 0.8 else len(args[1]) + 1
                if s and plen and rng.random() < 0.6 and plen <= len(s):
                    st = rng.randrange(0, len(s) - plen + 1)
                    args.append(s[st:st + plen])
                else:
                    args.append(rand_str(rng, plen, "ab"))
            elif k == "n":
                args.append(rng.randrange(0, 4))
            else:
                args.append(rng.randrange(0, 4))
        out.append((m, args, rng.randrange(0, 3)))
    return out


@pytest.mark.parametrize("fast", [0, 1], ids=["faithful", "fast"])
def test_random_cases_match_oracle(Graph, fast):
    """both recordings equal the reference algorithm on random small inputs, including empty strings and
    patterns, patterns longer than the string, repeated NULs inside the padding, |from| < |to|"""
    n_checked = 0
    for m, args, padding in random_cases(1234 + fast, 140):
        enc = encode_args(m, args, padding)
        kinds = SIGNATURES[m][0]
        if m in ("replace", "replacen") and len(enc[2]) > len(enc[1]) and (len(enc[0]) > 4 or len(enc[2]) > 2):
            continue  # handle_shorter_from is O(L * (|to| L + L)) char-ops: keep it tiny
        if not fast and m in ("repeat",) and len(enc[0]) > 3:
            continue
        try:
            ref = oracle_raw(m, enc)
        except P.FindTooLong:
            continue
        got, _ = run_method(Graph, m, enc, fast)
        rk = SIGNATURES[m][1]
        if rk == "u8":
            assert int(got) == int(ref), (m, args, padding)
        elif rk == "str":
            assert list(got) == list(ref), (m, args, padding)
        else:
            assert int(got[1]) == int(ref[1]), (m, args, padding)
            assert decode_result(m, got) == decode_result(m, ref), (m, args, padding)
        n_checked += 1
    assert n_checked > 80


def test_split_family_random_cases(Graph):
    """split family on random inputs, every buffer compared raw with the oracle: empty pattern, pattern absent,
    overlapping matches ("aaaa" / "aa"), n = 0, 1, larger than the number of pieces, 255; leading / trailing / repeated
    separators; strings with more than 15 matches (the counted form of the buffer index) and 8-9 char patterns (the
    folded far end of the blocking range in the depth-minimised scan)"""
    rng = random.Random(77)
    methods = sorted(k for k, v in SIGNATURES.items() if v[1] == "split")
    checked = 0
    for _ in range(140):
        m = rng.choice(methods)
        mode = rng.random()
        if mode < 0.6:
            s = "".join(rng.choice("ab. ") for _ in range(rng.randrange(0, 6)))
            pats = ["", ".", "a", "aa", "ab", " ", "b.", "aba"]
        elif mode < 0.85:
            s = "".join(rng.choice("ab") for _ in range(rng.randrange(8, 20)))
            pats = ["a", "ab", "aab", "abab", "b", ""]
        else:
            p = "".join(rng.choice("ab") for _ in range(rng.randrange(8, 10)))
            s = (p + "b") * 2 + "".join(rng.choice("ab") for _ in range(rng.randrange(0, 4)))
            pats = [p]
        args = [s]
        if m != "split_ascii_whitespace":
            args.append(rng.choice(pats))
        else:
            args = ["".join(rng.choice("ab \t\n") for _ in range(rng.randrange(0, 20)))]
        if m in ("splitn", "rsplitn"):
            args.append(rng.choice([0, 1, 2, 3, 4, 7, 17, 255]))
        enc = encode_args(m, args, rng.randrange(0, 3))
        ref = oracle_raw(m, enc)
        got, _ = run_method(Graph, m, enc, 1)
        assert [list(b) for b in got[0]] == [list(b) for b in ref[0]], (m, args)
        assert int(got[1]) == int(ref[1]), (m, args)
        checked += 1
    assert checked == 140


def test_splitn_with_a_clear_n(Graph):
    """splitn / rsplitn with n given in the clear (the CLI's *_clear forms, rsplit_once): the copy buffer index is
    decoded by one PBS per buffer with the cap folded into the table"""
    from plain_exec import blocks_of, run_program
    rng = random.Random(31)
    for _ in range(60):
        m = rng.choice(["splitn", "rsplitn"])
        s = "".join(rng.choice("ab. ") for _ in range(rng.randrange(0, 9)))
        pat = rng.choice(["", ".", "a", "aa", "ab", " ", "b."])
        n = rng.choice([0, 1, 2, 3, 4, 7, 17, 255])
        es, ep = [ord(c) for c in s] + [0] * rng.randrange(0, 3), [ord(c) for c in pat]
        g = Graph()
        slots, vals, ids = [], [], []
        for arr in (es, ep):
            i, sl = g.input_chars(len(arr))
            ids.append(i); slots.append(sl.reshape(-1)); vals.append(blocks_of(arr).reshape(-1))
        ids.append(g.trivial_chars([n]))
        bufs, found = g.split_op(m, ids, fast=True)
        g.mark_output([x for b in bufs for x in b] + [found])
        g.compile(1)
        plain = run_program(g, np.concatenate(slots), np.concatenate(vals))

        def char_val(cid):
            sl = g.char_slots([cid])[0]
            return sum(int(plain[int(sl[q])] % 16 & 3) << (2 * q) for q in range(4))

        got = [[char_val(c) for c in b] for b in bufs]
        ref = (P._split if m == "splitn" else P._rsplit)(list(es), list(ep), False, False, n)
        assert got == [list(b) for b in ref[0]] and char_val(found) == int(ref[1]), (m, s, pat, n)
        g.close()


def test_fast_split_scan_is_shallow(Graph):
    """the depth-minimised scan: one level per position instead of four (the reference's op order: 29 levels)"""
    enc = encode_args("split", ["hello", "ello"], 1)
    got_fast, info_fast = run_method(Graph, "split", enc, 1)
    got_ref, info_ref = run_method(Graph, "split", enc, 0)
    assert [list(b) for b in got_fast[0]] == [list(b) for b in got_ref[0]] == [list(b) for b in oracle_raw("split", enc)[0]]
    assert info_fast.n_levels <= 18 < info_ref.n_levels


def test_eq_and_comparisons_random_buffers(Graph):
    """== != < <= > >= on buffers of different sizes, equal prefixes, embedded NULs and non-NUL tails: the shallow
    recordings (no length computation; first difference and direction in one first-in PBS) against the oracle"""
    rng = random.Random(91)
    levels = {}
    for _ in range(220):
        m = rng.choice(["eq", "ne", "lt", "le", "gt", "ge", "eq_ignore_case"])
        la = rng.choice([0, 1, 2, 5, 14, 15, 16, 17, 31, 40])
        ea = [rng.choice([0, 65, 66, 97, 98, 200, 255]) if rng.random() < 0.15 else rng.choice([65, 66, 97, 98]) for _ in range(la)]
        mode = rng.random()
        if mode < 0.3:
            eb = list(ea)
        elif mode < 0.5:
            eb = list(ea[:rng.randrange(0, la + 1)])
        elif mode < 0.7 and la:
            k = rng.randrange(0, la)
            eb = ea[:k] + [rng.choice([65, 66, 97, 98, 0])] + ea[k + 1:]
        elif mode < 0.85:
            eb = ea + [rng.choice([65, 66, 0]) for _ in range(rng.randrange(1, 4))]
        else:
            eb = [rng.choice([65, 66, 97, 98]) for _ in range(rng.randrange(0, 20))]
        if rng.random() < 0.5:
            ea, eb = eb, ea
        ea, eb = ea + [0] * rng.randrange(0, 4), eb + [0] * rng.randrange(0, 4)
        ref = oracle_raw(m, [ea, eb])
        got, info = run_method(Graph, m, [ea, eb], 1)
        assert int(got) == int(ref), (m, ea, eb)
        levels[m] = max(levels.get(m, 0), info.n_levels)
    assert levels["eq"] <= 3 and levels["ne"] <= 3 and max(levels[k] for k in ("lt", "le", "gt", "ge")) <= 5, levels
    # config 3's shape: two 65-char buffers
    s = [rng.randrange(32, 127) for _ in range(64)] + [0]
    o = list(s)
    o[40] = 33 if s[40] != 33 else 34
    for m, fn in (("eq", P.eq), ("ge", P.ge), ("le", P.le)):
        got, info = run_method(Graph, m, [s, o], 1)
        assert int(got) == int(fn(s, o)) and info.n_levels <= (3 if m == "eq" else 5), (m, info.n_levels)
    # above 255 chars the u8 length of the reference can wrap: those sizes keep the length-based recording
    big = [97] * 257
    for m, fn in (("eq", P.eq), ("ge", P.ge)):
        got, _ = run_method(Graph, m, [big, big[:3]], 1)
        assert int(got) == int(fn(big, big[:3])), m


def test_fast_recording_is_shallow(Graph):
    """the point of the re-association: config 4 (contains over 257 chars, 8-char pattern) in a handful of levels"""
    rng = random.Random(4)
    s = [rng.randrange(97, 123) for _ in range(256)] + [0]
    pat = s[124:132]
    got_fast, info_fast = run_method(Graph, "contains", [s, pat], 1)
    assert got_fast == 1 == P.contains(s, pat)
    assert info_fast.n_levels <= 6 and info_fast.n_pbs < 10000
    got_find, info_find = run_method(Graph, "find", [s, pat], 1)
    assert got_find == 124 == P.find(s, pat)
    assert info_find.n_levels <= 8
    s2 = list(s)
    s2[130] = 65
    assert run_method(Graph, "contains", [s2, pat], 1)[0] == 0
    assert run_method(Graph, "find", [s2, pat], 1)[0] == 255


def test_faithful_contains_levels(Graph):
    """the reference's own op order: same value, serial depth"""
    s = P.encrypt_str("awesomezamaisawesome", 3)
    got, info = run_method(Graph, "contains", [s, [ord(c) for c in "zama"]], 0)
    assert got == 1
    assert info.n_levels > 15


def test_slot_alignment_for_sharding(Graph):
    """with slot_align = world every level's PBS results are one contiguous run padded to a multiple of world"""
    s = P.encrypt_str("hello world", 1)
    for world in (2, 8):
        kinds = "sp"
        got, info = run_method(Graph, "find", [s, [ord(c) for c in "wor"]], 1, slot_align=world)
        assert got == 6


def test_compaction_network_exhaustive(Graph):
    """bubble_zeroes_right as a routing network: all NUL patterns of length 7 + random longer ones"""
    for mask in range(128):
        s = [(65 + i) if (mask >> i & 1) else 0 for i in range(7)]
        got, _ = run_method(Graph, "concatenate", [s[:4], s[4:]], 1)
        assert list(got) == P.bubble_zeroes_right(s)
    rng = random.Random(9)
    for L in (16, 33, 70):
        s = [rng.choice([0, 0, rng.randrange(1, 128)]) for _ in range(L)]
        got, info = run_method(Graph, "concatenate", [s[:L // 2], s[L // 2:]], 1)
        assert list(got) == P.bubble_zeroes_right(s)


def test_random_expression_dags_of_char_primitives(Graph):
    """chains of the 11 FheAsciiChar primitives (results feeding later ops, flags mixed with u8 values, trivial
    constants in between): the value-set / noise bookkeeping of the graph must hold for compositions, not only for
    single ops"""
    ops2 = {"eq": P.c_eq, "ne": P.c_ne, "le": P.c_le, "lt": P.c_lt, "ge": P.c_ge, "gt": P.c_gt, "bitand": P.c_and,
            "bitor": P.c_or, "sub": P.c_sub, "add": P.c_add}
    ops1 = {"is_whitespace": P.c_is_whitespace, "is_uppercase": P.c_is_uppercase, "is_lowercase": P.c_is_lowercase, "flip": P.c_flip}
    rng = random.Random(2024)
    for trial in range(12):
        g = Graph()
        n_in = 6
        ids, slots = g.input_chars(n_in)
        vals = [rng.choice([0, 1, 2, 9, 32, 65, 90, 97, 122, 127, 128, 200, 255, rng.randrange(256)]) for _ in range(n_in)]
        pool = [(int(i), v) for i, v in zip(ids, vals)]
        for c in (0, 1, 32, 255):
            pool.append((int(g.trivial_chars([c])[0]), c))
        outs = []
        for step in range(40):
            kind = rng.random()
            if kind < 0.7:
                name = rng.choice(sorted(ops2))
                (a, va), (b, vb) = rng.choice(pool), rng.choice(pool)
                node, val = g.char_op(name, a, b), ops2[name](va, vb)
            elif kind < 0.85:
                name = rng.choice(sorted(ops1))
                a, va = rng.choice(pool)
                node, val = g.char_op(name, a), ops1[name](va)
            else:
                (c, vc), (t, vt), (f, vf) = rng.choice(pool), rng.choice(pool), rng.choice(pool)
                node, val = g.char_op("if_then_else", c, t, f), P.c_ite(vc, vt, vf)
            pool.append((node, val))
            outs.append((node, val))
        g.mark_output([n for n, _ in outs])
        g.compile()
        values = run_program(g, slots.reshape(-1), blocks_of(vals).reshape(-1))
        got = chars_of(values, g.char_slots([n for n, _ in outs]))
        assert [int(v) for v in got] == [v for _, v in outs], trial


# ---- operands that occur more than once: identical chars share one node, so a lazy sum can carry a
# coefficient m > 1 on ONE PBS output and its noise counts m^2 -- the recording must stay inside the budget
@pytest.mark.parametrize("s", ["h", "hi", "abc", "hell", "hello", "hello!"])
def test_repeat_clear_up_to_max_repetitions(Graph, s):
    for n in range(4, 17):                                   # the reference CLI allows n up to 16 (main.rs:17)
        enc = [[ord(ch) for ch in s] + [0], n]
        got, _ = run_method(Graph, "repeat_clear", enc, 1)
        assert list(got) == list(P.repeat_clear(*enc)), (s, n)


@pytest.mark.parametrize("method", ["lt", "le", "gt", "ge", "eq", "ne"])
def test_string_compared_with_itself(Graph, method):
    vals = [ord(ch) for ch in "hello"] + [0]
    g = Graph()
    ids, slots = g.input_chars(len(vals))
    _, rc = g.string_op(method, [ids, ids], fast=True)       # the SAME FheString on both sides
    g.mark_output([rc])
    g.compile(1)
    values = run_program(g, slots.reshape(-1), blocks_of(vals).reshape(-1))
    want = {"lt": 0, "le": 1, "gt": 0, "ge": 1, "eq": 1, "ne": 0}[method]
    assert int(chars_of(values, g.char_slots([rc]))[0]) == want


@pytest.mark.parametrize("n", [4, 7, 16])
def test_repeat_with_trivial_count(Graph, n):
    vals = [ord(ch) for ch in "ab"] + [0]
    g = Graph()
    ids, slots = g.input_chars(len(vals))
    nid = g.trivial_chars([n])
    rs, _ = g.string_op("repeat", [ids, nid], fast=True)
    g.mark_output(list(rs))
    g.compile(1)
    values = run_program(g, slots.reshape(-1), blocks_of(vals).reshape(-1))
    got = [int(v) for v in chars_of(values, g.char_slots(list(rs)))]
    assert got == list(P.repeat(vals, n))
