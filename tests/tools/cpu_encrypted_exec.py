"""Run a compiled job list on REAL ciphertexts at the real parameter set on the CPU: the oracle's keyswitch + f64-FFT
blind rotation stands in for the GPU kernels (same algorithm, same noise), leveled jobs are numpy u64 arithmetic.  Used
to check new graph recordings for noise, not only for value: the phase error of every PBS INPUT is measured against the
decoding margin (delta/2 = 2^58), which the plaintext interpretation (plain_exec.py) cannot see.

  python tests/tools/cpu_encrypted_exec.py ge 65 | eq 65 | compact 65 | split hello ello | rsplitn hello l 2 | replace hello ello _llo | contains STRING PATTERN
About 22 PBS/s per host core, so keep programs to a few thousand PBS."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from plain_exec import blocks_of, run_program  # noqa: E402


def execute(o, keys, fbsk, graph, in_slots, in_vals, seed=7):
    """returns (decrypted values of all slots, plaintext values, worst |phase error| over the PBS inputs in units of delta)"""
    info = graph.info
    jobs, off, npbs, first = graph.program()
    luts = np.stack([o.lut_poly(t) for t in graph.luts()]) if len(graph.luts()) else np.zeros((1, o.N), np.uint64)
    posts = [o.lut_post(t) for t in graph.luts()] if len(graph.luts()) else [0]
    arena = np.zeros((info.slots_used, o.big), np.uint64)
    arena[np.asarray(in_slots, np.int64)] = o.encrypt_big(keys, in_vals, seed=seed)
    tslots, tvals = graph.trivials()
    for s, v in zip(tslots, tvals):
        arena[int(s), -1] = np.uint64(int(v)) << np.uint64(o.delta_log)
    plain = run_program(graph, in_slots, in_vals)
    worst = 0.0
    delta = float(1 << o.delta_log)

    def combine(j):
        acc = np.zeros(o.big, np.uint64)
        with np.errstate(over="ignore"):
            for t in range(int(j["n_terms"])):
                acc += arena[int(j["src"][t])] * np.uint64(int(j["coeff"][t]) % (1 << 64))
            acc[-1] += np.uint64(int(j["constant"]))
        return acc

    for l in range(info.n_levels):
        a, b, n = int(off[l]), int(off[l + 1]), int(npbs[l])
        if n:
            lin = np.stack([combine(j) for j in jobs[a:a + n]])
            # noise check on the PBS inputs: phase - expected plaintext
            ph = o.phases(keys.s_glwe, lin)
            for k, j in enumerate(jobs[a:a + n]):
                want = int(j["constant"]) >> o.delta_log
                for t in range(int(j["n_terms"])):
                    want += int(j["coeff"][t]) * int(plain[int(j["src"][t])])
                err = (int(ph[k]) - ((want % 32) << o.delta_log)) % (1 << 64)
                if err >= 1 << 63:
                    err -= 1 << 64
                worst = max(worst, abs(err) / delta)
            out, _ = o.pbs_fft(keys, fbsk, luts, [int(j["lut"]) for j in jobs[a:a + n]], lin, post=posts)
            for k, j in enumerate(jobs[a:a + n]):
                arena[int(j["dst"])] = out[k]
        for j in jobs[a + n:b]:
            arena[int(j["dst"])] = combine(j)
    return o.decrypt_big(keys, arena).astype(np.int64), plain, worst


def main():
    from fhestring_b200.graph import Graph
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    what = sys.argv[1]
    t0 = time.time()
    o = Oracle(**P)
    keys = o.keygen(1)
    fbsk = o.fourier_bsk(keys)
    print(f"keys {time.time() - t0:.1f} s", flush=True)
    rng = np.random.default_rng(3)
    g = Graph()
    ids, slots, vals = [], [], []

    def add(chars):
        i, s = g.input_chars(len(chars))
        ids.append(i); slots.append(s.reshape(-1)); vals.append(blocks_of(list(chars)).reshape(-1))

    if what in ("ge", "eq", "le", "lt", "gt", "ne"):
        L = int(sys.argv[2])
        a = [int(x) for x in rng.integers(32, 127, L - 1)] + [0]
        b = list(a)
        b[(2 * L) // 3] = a[(2 * L) // 3] ^ 1
        add(a); add(b)
        rs, rc = g.string_op(what, ids, fast=True)
        outs = [rc]
    elif what in ("replace", "contains", "find"):
        add([ord(c) for c in sys.argv[2]] + [0])
        for extra in sys.argv[3:]:
            add([ord(c) for c in extra])
        rs, rc = g.string_op(what, ids, fast=True)
        outs = (list(rs) if rs is not None else []) + ([rc] if rc is not None else [])
    elif what == "compact":
        L = int(sys.argv[2])
        s = [int(x) if rng.random() < 0.6 else 0 for x in rng.integers(1, 256, L)]
        add(s[:L // 2]); add(s[L // 2:])
        rs, rc = g.string_op("concatenate", ids, fast=True)
        outs = list(rs)
    else:
        add([ord(c) for c in sys.argv[2]] + [0]); add([ord(c) for c in sys.argv[3]])
        if len(sys.argv) > 4:   # the n of splitn / rsplitn: one encrypted u8
            i, s_ = g.input_chars(1)
            ids.append(i); slots.append(s_.reshape(-1)); vals.append(blocks_of([int(sys.argv[4])]).reshape(-1))
        bufs, found = g.split_op(what, ids, fast=True)
        outs = [x for bb in bufs for x in bb] + [found]
    g.mark_output(outs)
    info = g.compile(1)
    print(f"{what}: {info.n_pbs} PBS in {info.n_levels} levels, {info.slots_used} slots", flush=True)
    t1 = time.time()
    got, plain, worst = execute(o, keys, fbsk, g, np.concatenate(slots), np.concatenate(vals))
    jobs, off, npbs, first = g.program()
    bad = [int(j["dst"]) for j in jobs if got[int(j["dst"])] != plain[int(j["dst"])] % 16]
    print(f"executed in {time.time() - t1:.1f} s; blocks written {len(jobs)}, wrong {len(bad)}; "
          f"worst |phase error| of a PBS input = {worst:.4f} delta (margin 0.5)")
    assert not bad and worst < 0.25


if __name__ == "__main__":
    main()
