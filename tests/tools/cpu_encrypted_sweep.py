"""Every reference case (the reference's 43 unit tests + the CLI self-check inputs, tests/golden/reference_tests.json)
recorded in fast mode and executed on REAL ciphertexts at the real parameter set on the CPU (cpu_encrypted_exec.execute):
each written block must decrypt to the plaintext interpretation; the worst phase error over all PBS inputs is printed.
  python tests/tools/cpu_encrypted_sweep.py run [max PBS per case, default all]     (about 77 PBS/s on 8 cores)
  python tests/tools/cpu_encrypted_sweep.py dry                                     (sizes only)"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import numpy as np
from strcases import SIGNATURES, encode_args, reference_cases
from plain_exec import blocks_of
from cpu_encrypted_exec import execute
from fhestring_b200.graph import Graph
from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
dry = len(sys.argv) > 1 and sys.argv[1] == "dry"
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 10**9
o = Oracle(**P); keys = o.keygen(1); fbsk = o.fourier_bsk(keys)
tot = 0; worst_all = 0; t0 = time.time(); done = 0
for c in reference_cases():
    m = c["method"]
    if isinstance(c["expect"], str) and c["expect"].startswith("panic"): continue
    kinds, rk = SIGNATURES[m]
    enc = encode_args(m, c["args"], c["padding"])
    g = Graph(); ids, slots, vals, clear_n = [], [], [], 0
    for kind, v in zip(kinds, enc):
        if kind == "c": clear_n = v; continue
        v = [v] if kind == "n" else list(v)
        i, s = g.input_chars(len(v)); ids.append(i); slots.append(s.reshape(-1)); vals.append(blocks_of(v).reshape(-1))
    if rk == "split":
        bufs, found = g.split_op(m, ids, fast=True); outs = [x for b in bufs for x in b] + [found]
    else:
        rs, rc = g.string_op(m, ids, fast=True, clear_n=clear_n)
        outs = (list(rs) if rs is not None else []) + ([rc] if rc is not None else [])
    g.mark_output(outs); info = g.compile(1)
    tot += info.n_pbs
    if dry or info.n_pbs > limit or info.n_pbs == 0:
        g.close(); continue
    got, plain, worst = execute(o, keys, fbsk, g, np.concatenate(slots), np.concatenate(vals))
    jobs, off, npbs, first = g.program()
    bad = [int(j["dst"]) for j in jobs if got[int(j["dst"])] != plain[int(j["dst"])] % 16]
    worst_all = max(worst_all, worst); done += 1
    print(f"{c['name']:28s} pbs {info.n_pbs:6d} levels {info.n_levels:3d} wrong {len(bad)} worst {worst:.4f}", flush=True)
    assert not bad
    g.close()
print("total pbs", tot, "cases run", done, "worst", worst_all, "time", round(time.time() - t0), "s")
