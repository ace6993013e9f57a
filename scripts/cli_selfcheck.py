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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


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
    from cli_config1 import run_all          # the method table lives with the -m gpu test that runs it under the driver
    from fhestring_b200.fhestring import MyClientKey

    ck = MyClientKey.from_params(seed=7)
    sk = ck.get_server_key(arena_blocks=1 << 18, fast=not a.faithful)

    def show(r):
        print(f"{'Test Passed' if r['passed'] else 'Test Failed'}  {r['method']:24s} {r['ms']:9.1f} ms   (last flush: {r['pbs']} PBS in "
              f"{r['levels']} levels)" + ("" if r["passed"] else f"   std {r['std']!r} oracle {r['oracle']!r} got {r['got']!r}"), flush=True)

    rows = run_all(ck, sk, ck.get_public_parameters(), a.string, a.pattern, a.n, a.frm, a.to, on_row=show)
    failed = sum(not r["passed"] for r in rows)
    print(f"{len(rows) - failed} of {len(rows)} methods passed; total {sum(r['ms'] for r in rows) / 1e3:.2f} s")
    if a.json:
        with open(a.json, "w") as fj:
            json.dump(dict(args=vars(a), rows=rows), fj, indent=1, default=str)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
