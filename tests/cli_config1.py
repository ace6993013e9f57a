"""BASELINE config 1: the reference CLI's self-check loop (/root/reference/src/main.rs:34-116 dispatching to
src/utils.rs:122-718) through the host API of this repo -- the same 52 methods in the same order, inputs re-encrypted
per method, result decrypted and compared with (i) what Rust `std` gives (spelled out in Python) and (ii) the
plaintext restatement of the reference's own algorithm (oracle/fhestring_plain.py).  Used by the `-m gpu` test
tests/test_gpu_strings.py::test_config1_cli_all_52_methods and by scripts/cli_selfcheck.py."""
from __future__ import annotations

import time

from oracle import fhestring_plain as P
from strcases import SIGNATURES, decode_result, encode_args

WS = " \t\n\r\x0b\x0c"
STRING_PADDING = 1          # main.rs:12
MAX_FIND_LENGTH = 255       # main.rs:20

METHODS = ("Contains ContainsClear EndsWith EndsWithClear EqIgnoreCase Find FindClear IsEmpty Len Repeat RepeatClear Replace "
           "ReplaceClear ReplaceN ReplaceNClear Rfind RfindClear Rsplit RsplitClear RsplitOnce RsplitOnceClear RsplitN "
           "RsplitNClear RsplitTerminator RsplitTerminatorClear Split SplitClear SplitAsciiWhitespace SplitInclusive "
           "SplitInclusiveClear SplitTerminator SplitTerminatorClear SplitN SplitNClear StartsWith StartsWithClear StripPrefix "
           "StripPrefixClear StripSuffix StripSuffixClear ToLower ToUpper Trim TrimEnd TrimStart Concatenate Lt Le Gt Ge Eq Ne").split()
assert len(METHODS) == 52

SNAKE = {"EqIgnoreCase": "eq_ignore_case", "IsEmpty": "is_empty", "ReplaceN": "replacen", "RsplitOnce": "rsplit_once",
         "RsplitN": "rsplitn", "RsplitTerminator": "rsplit_terminator", "SplitAsciiWhitespace": "split_ascii_whitespace",
         "SplitInclusive": "split_inclusive", "SplitTerminator": "split_terminator", "SplitN": "splitn",
         "StartsWith": "starts_with", "EndsWith": "ends_with", "StripPrefix": "strip_prefix", "StripSuffix": "strip_suffix",
         "ToLower": "to_lower", "ToUpper": "to_upper", "TrimEnd": "trim_end", "TrimStart": "trim_start",
         "RepeatClear": "repeat_clear"}


def base_method(m: str) -> str:
    """the oracle / signature name behind a CLI method: the *_clear forms delegate to the encrypted-pattern form with a
    trivially encrypted pattern (mod.rs:198-211), so they share its algorithm"""
    if m == "RepeatClear":
        return "repeat_clear"
    b = m[:-5] if m.endswith("Clear") else m
    return SNAKE.get(b, b.lower())


def trimv(v):
    v = list(v)
    while v and v[0] == "":
        v.pop(0)
    while v and v[-1] == "":
        v.pop()
    return v


def split_terminator(s, p):
    parts = s.split(p)
    return parts[:-1] if parts and parts[-1] == "" else parts


def split_inclusive(s, p):
    parts = s.split(p)
    return [x + p for x in parts[:-1]] + ([parts[-1]] if parts[-1] else [])


def oracle_result(m, h, p, n, f, t):
    """decoded result of the reference's algorithm on the same plaintext arguments"""
    b = base_method(m)
    kinds = SIGNATURES[b][0]
    if b in ("replace", "replacen"):
        args = [h, f, t] + ([n] if b == "replacen" else [])
    elif kinds == "s":
        args = [h]
    elif kinds in ("sn", "sc"):
        args = [h, n]
    elif kinds == "spn":
        args = [h, p, n]
    else:
        args = [h, p]
    fn = getattr(P, "length" if b == "len" else b)
    return decode_result(b, fn(*encode_args(b, args, STRING_PADDING)))


def run_all(ck, sk, pp, h, p, n, f, t, on_row=None):
    """-> rows [{method, passed, got, std, oracle, ms, pbs, levels}] for the 52 methods, in the CLI's order"""
    from fhestring_b200.fhestring import FheSplit, FheStrip
    assert n <= 16, "n must be <= MAX_REPETITIONS"   # main.rs:37-40

    def split_res(r):
        bufs, found = FheSplit.decrypt(r, ck)
        return trimv(bufs)

    def strip_res(r, expect_found, expect_str):
        s, found = FheStrip.decrypt(r, ck)
        return [s, found], [expect_str if expect_found else h, int(expect_found)]

    rows = []
    for m in METHODS:
        sk.reset()
        t0 = time.perf_counter()
        S = ck.encrypt(h, STRING_PADDING, pp, sk.key)
        Pn = ck.encrypt_no_padding(p)
        F, T, N = ck.encrypt_no_padding(f), ck.encrypt_no_padding(t), ck.encrypt_char(n)
        find_exp = lambda i: i if i >= 0 else MAX_FIND_LENGTH
        if m == "Contains": got, exp = ck.decrypt_char(sk.contains(S, Pn, pp)), int(p in h)
        elif m == "ContainsClear": got, exp = ck.decrypt_char(sk.contains_clear(S, p, pp)), int(p in h)
        elif m == "EndsWith": got, exp = ck.decrypt_char(sk.ends_with(S, Pn, pp)), int(h.endswith(p))
        elif m == "EndsWithClear": got, exp = ck.decrypt_char(sk.ends_with_clear(S, p, pp)), int(h.endswith(p))
        elif m == "EqIgnoreCase": got, exp = ck.decrypt_char(sk.eq_ignore_case(S, ck.encrypt(p, STRING_PADDING, pp, sk.key), pp)), int(h.lower() == p.lower())
        elif m == "Find": got, exp = ck.decrypt_char(sk.find(S, Pn, pp)), find_exp(h.find(p))
        elif m == "FindClear": got, exp = ck.decrypt_char(sk.find_clear(S, p, pp)), find_exp(h.find(p))
        elif m == "IsEmpty": got, exp = ck.decrypt_char(sk.is_empty(S, pp)), int(h == "")
        elif m == "Len": got, exp = ck.decrypt_char(sk.len(S, pp)), len(h)
        elif m == "Repeat": got, exp = ck.decrypt(sk.repeat(S, N, pp)), h * n
        elif m == "RepeatClear": got, exp = ck.decrypt(sk.repeat_clear(S, n, pp)), h * n
        elif m == "Replace": got, exp = ck.decrypt(sk.replace(S, F, T, pp)), h.replace(f, t)
        elif m == "ReplaceClear": got, exp = ck.decrypt(sk.replace_clear(S, f, t, pp)), h.replace(f, t)
        elif m == "ReplaceN": got, exp = ck.decrypt(sk.replacen(S, F, T, N, pp)), h.replace(f, t, n)
        elif m == "ReplaceNClear": got, exp = ck.decrypt(sk.replacen_clear(S, f, t, n, pp)), h.replace(f, t, n)
        elif m == "Rfind": got, exp = ck.decrypt_char(sk.rfind(S, Pn, pp)), find_exp(h.rfind(p))
        elif m == "RfindClear": got, exp = ck.decrypt_char(sk.rfind_clear(S, p, pp)), find_exp(h.rfind(p))
        elif m == "Rsplit": got, exp = split_res(sk.rsplit(S, Pn, pp)), trimv(h.split(p)[::-1])
        elif m == "RsplitClear": got, exp = split_res(sk.rsplit_clear(S, p, pp)), trimv(h.split(p)[::-1])
        elif m in ("RsplitOnce", "RsplitOnceClear"):
            r = sk.rsplit_once(S, Pn, pp) if m == "RsplitOnce" else sk.rsplit_once_clear(S, p, pp)
            got = split_res(r)
            exp = trimv([h.rsplit(p, 1)[1], h.rsplit(p, 1)[0]]) if p in h else got   # utils.rs: only compared when std finds it
        elif m == "RsplitN": got, exp = split_res(sk.rsplitn(S, Pn, N, pp)), trimv(h.rsplit(p, n - 1)[::-1] if n else [])
        elif m == "RsplitNClear": got, exp = split_res(sk.rsplitn_clear(S, p, n, pp)), trimv(h.rsplit(p, n - 1)[::-1] if n else [])
        elif m == "RsplitTerminator": got, exp = split_res(sk.rsplit_terminator(S, Pn, pp)), trimv(split_terminator(h, p)[::-1])
        elif m == "RsplitTerminatorClear": got, exp = split_res(sk.rsplit_terminator_clear(S, p, pp)), trimv(split_terminator(h, p)[::-1])
        elif m == "Split": got, exp = split_res(sk.split(S, Pn, pp)), trimv(h.split(p))
        elif m == "SplitClear": got, exp = split_res(sk.split_clear(S, p, pp)), trimv(h.split(p))
        elif m == "SplitAsciiWhitespace": got, exp = split_res(sk.split_ascii_whitespace(S, pp)), trimv(h.split())
        elif m == "SplitInclusive": got, exp = split_res(sk.split_inclusive(S, Pn, pp)), trimv(split_inclusive(h, p))
        elif m == "SplitInclusiveClear": got, exp = split_res(sk.split_inclusive_clear(S, p, pp)), trimv(split_inclusive(h, p))
        elif m == "SplitTerminator": got, exp = split_res(sk.split_terminator(S, Pn, pp)), trimv(split_terminator(h, p))
        elif m == "SplitTerminatorClear": got, exp = split_res(sk.split_terminator_clear(S, p, pp)), trimv(split_terminator(h, p))
        elif m == "SplitN": got, exp = split_res(sk.splitn(S, Pn, N, pp)), trimv(h.split(p, n - 1) if n else [])
        elif m == "SplitNClear": got, exp = split_res(sk.splitn_clear(S, p, n, pp)), trimv(h.split(p, n - 1) if n else [])
        elif m == "StartsWith": got, exp = ck.decrypt_char(sk.starts_with(S, Pn, pp)), int(h.startswith(p))
        elif m == "StartsWithClear": got, exp = ck.decrypt_char(sk.starts_with_clear(S, p, pp)), int(h.startswith(p))
        elif m == "StripPrefix": got, exp = strip_res(sk.strip_prefix(S, Pn, pp), h.startswith(p), h[len(p):])
        elif m == "StripPrefixClear": got, exp = strip_res(sk.strip_prefix_clear(S, p, pp), h.startswith(p), h[len(p):])
        elif m == "StripSuffix": got, exp = strip_res(sk.strip_suffix(S, Pn, pp), h.endswith(p), h[:len(h) - len(p)])
        elif m == "StripSuffixClear": got, exp = strip_res(sk.strip_suffix_clear(S, p, pp), h.endswith(p), h[:len(h) - len(p)])
        elif m == "ToLower": got, exp = ck.decrypt(sk.to_lower(S, pp)), h.lower()
        elif m == "ToUpper": got, exp = ck.decrypt(sk.to_upper(S, pp)), h.upper()
        elif m == "Trim": got, exp = ck.decrypt(sk.trim(S, pp)), h.strip(WS)
        elif m == "TrimEnd": got, exp = ck.decrypt(sk.trim_end(S, pp)), h.rstrip(WS)
        elif m == "TrimStart": got, exp = ck.decrypt(sk.trim_start(S, pp)), h.lstrip(WS)
        elif m == "Concatenate": got, exp = ck.decrypt(sk.concatenate(S, ck.encrypt(p, STRING_PADDING, pp, sk.key), pp)), h + p
        else:
            O = ck.encrypt(p, STRING_PADDING, pp, sk.key)
            fn = {"Lt": sk.lt, "Le": sk.le, "Gt": sk.gt, "Ge": sk.ge, "Eq": sk.eq, "Ne": sk.ne}[m]
            exp = int({"Lt": h < p, "Le": h <= p, "Gt": h > p, "Ge": h >= p, "Eq": h == p, "Ne": h != p}[m])
            got = ck.decrypt_char(fn(S, O, pp))
        dt = time.perf_counter() - t0
        orc = oracle_result(m, h, p, n, f, t)
        ok_std = got == exp
        ok_orc = got == orc
        info = sk.last_info
        row = dict(method=m, ms=1e3 * dt, passed=bool(ok_std and ok_orc), passed_std=bool(ok_std), passed_oracle=bool(ok_orc),
                   got=got, std=exp, oracle=orc, pbs=int(info.n_pbs) if info else 0, levels=int(info.n_levels) if info else 0)
        rows.append(row)
        if on_row:
            on_row(row)
    return rows
