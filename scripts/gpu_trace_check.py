"""Debug aid (GPU): run one string method through the graph + engine and compare EVERY arena block the program
writes with the plaintext interpretation of the same job list (tests/plain_exec.py).  Prints the first
mismatching jobs.  usage: python scripts/gpu_trace_check.py METHOD ARG... [--padding K] [--faithful] [--repeat R]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from plain_exec import blocks_of, run_program  # noqa: E402
from strcases import SIGNATURES, encode_args  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("method", help="a method name, or ALL for every reference unit-test case in sequence on one engine")
    ap.add_argument("args", nargs="*")
    ap.add_argument("--padding", type=int, default=1)
    ap.add_argument("--faithful", action="store_true")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--n", type=int, default=742, help="small LWE dimension (use e.g. 16 under compute-sanitizer)")
    a = ap.parse_args()
    from fhestring_b200.client import ClientKey
    from fhestring_b200.engine import Engine
    from fhestring_b200.graph import Graph

    ck = ClientKey(seed=20, n=a.n)
    bsk, ksk = ck.server_keys()
    eng = Engine(arena_blocks=1 << 17, n=a.n)
    eng.load_keys(bsk, ksk)
    if a.method == "ALL":
        from strcases import reference_cases
        work = [(c["method"], encode_args(c["method"], c["args"], c["padding"]), c["name"]) for c in reference_cases()
                if not (isinstance(c["expect"], str) and c["expect"].startswith("panic"))
                and SIGNATURES[c["method"]][1] != "split" and c["name"] not in ("replace2", "replacen")] * a.repeat
    else:
        kinds = SIGNATURES[a.method][0]
        work = [(a.method, encode_args(a.method, [int(x) if k in "nc" else x for k, x in zip(kinds, a.args)], a.padding), a.method)] * a.repeat
    for rep, (method, enc, name) in enumerate(work):
        a.method = method
        kinds = SIGNATURES[method][0]
        g = Graph()
        ids, slots, vals, clear_n = [], [], [], 0
        for kind, v in zip(kinds, enc):
            if kind == "c":
                clear_n = v
                continue
            v = [v] if kind == "n" else list(v)
            i, s = g.input_chars(len(v))
            ids.append(i); slots.append(s.reshape(-1)); vals.append(blocks_of(v).reshape(-1))
        rs, rc = g.string_op(a.method, ids, fast=not a.faithful, clear_n=clear_n)
        outs = ([] if rs is None else list(rs)) + ([] if rc is None else [rc])
        g.mark_output(outs)
        info = g.compile(1)
        in_slots, in_vals = np.concatenate(slots), np.concatenate(vals)
        eng.upload(int(in_slots[0]), ck.encrypt_blocks(in_vals.astype(np.uint8)))
        plain = run_program(g, in_slots, in_vals)
        jobs, off, npbs, first = g.program()
        g.execute(eng)
        got = ck.decrypt_blocks(eng.download(0, info.slots_used)).astype(np.int64)
        bad = [(l, int(j["dst"])) for l in range(info.n_levels) for j in jobs[off[l]:off[l + 1]]
               if got[int(j["dst"])] != plain[int(j["dst"])] % 16]
        if bad or not a.quiet:
            print(f"{name} {rep}: levels {info.n_levels} pbs {info.n_pbs} slots {info.slots_used} mismatching blocks {len(bad)}", flush=True)
        total_pbs = locals().get("total_pbs", 0) + int(info.n_pbs)
        if bad:   # is it reproducible? run the same program again on the same inputs
            eng.upload(int(in_slots[0]), ck.encrypt_blocks(in_vals.astype(np.uint8)))
            g2 = None
            prog = None
            print("  phase errors of the mismatching blocks (units of 2^-64):", [int(e) for e in ck.decrypt_blocks(eng.download(bad[0][1], 1), with_error=True)[1]])
        for l, d in bad[:10]:
            j = [x for x in jobs if int(x["dst"]) == d][0]
            nt = int(j["n_terms"])
            print(f"  level {l} dst {d} lut {int(j['lut'])} got {got[d]} want {plain[d] % 16} terms",
                  [(int(j['src'][t]), int(j['coeff'][t]), int(got[int(j['src'][t])]), int(plain[int(j['src'][t])])) for t in range(nt)],
                  "const", int(j["constant"]) >> 59)
        eng_lut_count = len(g.luts())
        g.close()
    print("done; PBS executed:", locals().get("total_pbs", 0))
    eng.close()


if __name__ == "__main__":
    main()
