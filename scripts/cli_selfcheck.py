#!/usr/bin/env python3
"""BASELINE config 1 on the GPU: the reference CLI's self-check loop (/root/reference/src/main.rs:34-116,
src/utils.rs:122-718) through the host API of this repo -- same 52 methods in the same order, inputs re-encrypted per
method, result decrypted and compared with what Rust `std` gives (spelled out in Python), wall time per method.

  python scripts/cli_selfcheck.py --string hello --pattern ello --n 1 --from ello --to _llo [--faithful]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WS = " \t\n\r\x0b\x0c"
STRING_PADDING = 1          # main.rs:12
MAX_FIND_LENGTH = 255       # main.rs:20

METHODS = ("Contains ContainsClear EndsWith EndsWithClear EqIgnoreCase Find FindClear IsEmpty Len Repeat RepeatClear Replace "
           "ReplaceClear ReplaceN ReplaceNClear Rfind RfindClear Rsplit RsplitClear RsplitOnce RsplitOnceClear RsplitN "
           "RsplitNClear RsplitTerminator RsplitTerminatorClear Split SplitClear SplitAsciiWhitespace SplitInclusive "
           "SplitInclusiveClear SplitTerminator SplitTerminatorClear SplitN SplitNClear StartsWith StartsWithClear StripPrefix "
           "StripPrefixClear StripSuffix StripSuffixClear ToLower ToUpper Trim TrimEnd TrimStart Concatenate Lt Le Gt Ge Eq Ne").split()


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--string", default="hello")
    ap.add_argument("--pattern", default="ello")
    ap.add_argument("--n", type=int, default=1)
    ap.add_argument("--from", dest="frm", default="ello")
    ap.add_argument("--to", default="_llo")
    ap.add_argument("--faithful", action="store_true", help="record the reference's own op order instead of the depth-minimised one")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    assert a.n <= 16, "n must be <= MAX_REPETITIONS"   # main.rs:37-40
    from fhestring_b200.fhestring import FheSplit, FheStrip, MyClientKey

    ck = MyClientKey.from_params(seed=7)
    sk = ck.get_server_key(arena_blocks=1 << 18, fast=not a.faithful)
    pp = ck.get_public_parameters()
    h, p, n, f, t = a.string, a.pattern, a.n, a.frm, a.to

    def split_res(r):
        bufs, found = FheSplit.decrypt(r, ck)
        return trimv(bufs)

    def strip_res(r, expect_found, expect_str):
        s, found = FheStrip.decrypt(r, ck)
        return (found, s if found else None), (int(expect_found), expect_str if expect_found else None)

    rows, failed = [], 0
    for m in METHODS:
        sk.reset()
        t0 = time.perf_counter()
        S = ck.encrypt(h, STRING_PADDING, pp, sk.key)
        P = ck.encrypt_no_padding(p)
        F, T, N = ck.encrypt_no_padding(f), ck.encrypt_no_padding(t), ck.encrypt_char(n)
        find_exp = lambda i: i if i >= 0 else MAX_FIND_LENGTH
        if m == "Contains": got, exp = ck.decrypt_char(sk.contains(S, P, pp)), int(p in h)
        elif m == "ContainsClear": got, exp = ck.decrypt_char(sk.contains_clear(S, p, pp)), int(p in h)
        elif m == "EndsWith": got, exp = ck.decrypt_char(sk.ends_with(S, P, pp)), int(h.endswith(p))
        elif m == "EndsWithClear": got, exp = ck.decrypt_char(sk.ends_with_clear(S, p, pp)), int(h.endswith(p))
        elif m == "EqIgnoreCase": got, exp = ck.decrypt_char(sk.eq_ignore_case(S, ck.encrypt(p, STRING_PADDING, pp, sk.key), pp)), int(h.lower() == p.lower())
        elif m == "Find": got, exp = ck.decrypt_char(sk.find(S, P, pp)), find_exp(h.find(p))
        elif m == "FindClear": got, exp = ck.decrypt_char(sk.find_clear(S, p, pp)), find_exp(h.find(p))
        elif m == "IsEmpty": got, exp = ck.decrypt_char(sk.is_empty(S, pp)), int(h == "")
        elif m == "Len": got, exp = ck.decrypt_char(sk.len(S, pp)), len(h)
        elif m == "Repeat": got, exp = ck.decrypt(sk.repeat(S, N, pp)), h * n
        elif m == "RepeatClear": got, exp = ck.decrypt(sk.repeat_clear(S, n, pp)), h * n
        elif m == "Replace": got, exp = ck.decrypt(sk.replace(S, F, T, pp)), h.replace(f, t)
        elif m == "ReplaceClear": got, exp = ck.decrypt(sk.replace_clear(S, f, t, pp)), h.replace(f, t)
        elif m == "ReplaceN": got, exp = ck.decrypt(sk.replacen(S, F, T, N, pp)), h.replace(f, t, n)
        elif m == "ReplaceNClear": got, exp = ck.decrypt(sk.replacen_clear(S, f, t, n, pp)), h.replace(f, t, n)
        elif m == "Rfind": got, exp = ck.decrypt_char(sk.rfind(S, P, pp)), find_exp(h.rfind(p))
        elif m == "RfindClear": got, exp = ck.decrypt_char(sk.rfind_clear(S, p, pp)), find_exp(h.rfind(p))
        elif m == "Rsplit": got, exp = split_res(sk.rsplit(S, P, pp)), trimv(h.split(p)[::-1])
        elif m == "RsplitClear": got, exp = split_res(sk.rsplit_clear(S, p, pp)), trimv(h.split(p)[::-1])
        elif m in ("RsplitOnce", "RsplitOnceClear"):
            r = sk.rsplit_once(S, P, pp) if m == "RsplitOnce" else sk.rsplit_once_clear(S, p, pp)
            got = split_res(r)
            exp = trimv([h.rsplit(p, 1)[1], h.rsplit(p, 1)[0]]) if p in h else got   # utils.rs: only compared when std finds it
        elif m == "RsplitN": got, exp = split_res(sk.rsplitn(S, P, N, pp)), trimv(h.rsplit(p, n - 1)[::-1] if n else [])
        elif m == "RsplitNClear": got, exp = split_res(sk.rsplitn_clear(S, p, n, pp)), trimv(h.rsplit(p, n - 1)[::-1] if n else [])
        elif m == "RsplitTerminator": got, exp = split_res(sk.rsplit_terminator(S, P, pp)), trimv(split_terminator(h, p)[::-1])
        elif m == "RsplitTerminatorClear": got, exp = split_res(sk.rsplit_terminator_clear(S, p, pp)), trimv(split_terminator(h, p)[::-1])
        elif m == "Split": got, exp = split_res(sk.split(S, P, pp)), trimv(h.split(p))
        elif m == "SplitClear": got, exp = split_res(sk.split_clear(S, p, pp)), trimv(h.split(p))
        elif m == "SplitAsciiWhitespace": got, exp = split_res(sk.split_ascii_whitespace(S, pp)), trimv(h.split())
        elif m == "SplitInclusive": got, exp = split_res(sk.split_inclusive(S, P, pp)), trimv(split_inclusive(h, p))
        elif m == "SplitInclusiveClear": got, exp = split_res(sk.split_inclusive_clear(S, p, pp)), trimv(split_inclusive(h, p))
        elif m == "SplitTerminator": got, exp = split_res(sk.split_terminator(S, P, pp)), trimv(split_terminator(h, p))
        elif m == "SplitTerminatorClear": got, exp = split_res(sk.split_terminator_clear(S, p, pp)), trimv(split_terminator(h, p))
        elif m == "SplitN": got, exp = split_res(sk.splitn(S, P, N, pp)), trimv(h.split(p, n - 1) if n else [])
        elif m == "SplitNClear": got, exp = split_res(sk.splitn_clear(S, p, n, pp)), trimv(h.split(p, n - 1) if n else [])
        elif m == "StartsWith": got, exp = ck.decrypt_char(sk.starts_with(S, P, pp)), int(h.startswith(p))
        elif m == "StartsWithClear": got, exp = ck.decrypt_char(sk.starts_with_clear(S, p, pp)), int(h.startswith(p))
        elif m == "StripPrefix": got, exp = strip_res(sk.strip_prefix(S, P, pp), h.startswith(p), h[len(p):])
        elif m == "StripPrefixClear": got, exp = strip_res(sk.strip_prefix_clear(S, p, pp), h.startswith(p), h[len(p):])
        elif m == "StripSuffix": got, exp = strip_res(sk.strip_suffix(S, P, pp), h.endswith(p), h[:len(h) - len(p)])
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
        ok = got == exp
        failed += not ok
        info = sk.last_info
        rows.append(dict(method=m, ms=1e3 * dt, passed=bool(ok), pbs=int(info.n_pbs) if info else 0, levels=int(info.n_levels) if info else 0))
        print(f"{'Test Passed' if ok else 'Test Failed'}  {m:24s} {1e3 * dt:9.1f} ms   (last flush: {rows[-1]['pbs']} PBS in {rows[-1]['levels']} levels)"
              + ("" if ok else f"   expected {exp!r} got {got!r}"), flush=True)
    print(f"{len(rows) - failed} of {len(rows)} methods passed; total {sum(r['ms'] for r in rows) / 1e3:.2f} s")
    if a.json:
        with open(a.json, "w") as fj:
            json.dump(dict(args=vars(a), rows=rows), fj, indent=1)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
