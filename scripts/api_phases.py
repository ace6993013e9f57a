"""Where the host time of one API-level string query goes (MyServerKey.<method>(...) + download): adopt (stack + upload),
record, compile, execute, download -- the gap between `api_latency_ms` and `latency_ms` of bench.py.
usage (GPU box): python scripts/api_phases.py  -> gpurun_out/api_phases.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhestring_b200.client import ClientKey  # noqa: E402
from fhestring_b200.engine import Engine  # noqa: E402
from fhestring_b200.fhestring import FheAsciiChar, FheString, MyServerKey  # noqa: E402

ck = ClientKey(seed=1)
bsk, ksk = ck.server_keys()
eng = Engine(arena_blocks=1 << 16)
eng.load_keys(bsk, ksk)
sk = MyServerKey(None, None, engine=eng)
rng = np.random.default_rng(3)


def text(n):
    return rng.integers(97, 123, n).astype(np.uint8)


def run(method, vals, reps=4):
    cts_np = ck.encrypt_u8(np.concatenate(vals)).reshape(-1, eng.big)
    best = None
    for _ in range(reps):
        sk.reset()
        ph = {}
        t0 = time.perf_counter()
        off, args = 0, []
        for i, v in enumerate(vals):
            chars = [FheAsciiChar(ct=cts_np[4 * (off + j):4 * (off + j + 1)]) for j in range(len(v))]
            args.append(FheString(chars) if i == 0 or method in ("eq", "ge", "le") else chars)
            off += len(v)
        ph["wrap"] = time.perf_counter() - t0
        t = time.perf_counter()
        ids = [sk._ids(a) for a in args]
        ph["adopt (stack + upload)"] = time.perf_counter() - t
        t = time.perf_counter()
        rs, rc = sk.graph.string_op(method, ids, fast=True)
        ph["record"] = time.perf_counter() - t
        outs = [sk._wrap(i) for i in rs] if rs is not None and len(rs) else [sk._wrap(rc)]
        t = time.perf_counter()
        oid = np.array([c.id for c in outs], np.uint32)
        sk.graph.mark_output(oid)
        info = sk.graph.compile(1)
        ph["compile"] = time.perf_counter() - t
        t = time.perf_counter()
        sk.graph.execute(eng, 0, 1)
        eng.sync()
        ph["execute (bind + program upload + run)"] = time.perf_counter() - t
        t = time.perf_counter()
        raw = sk._download(outs)
        ph["download"] = time.perf_counter() - t
        ph["total"] = time.perf_counter() - t0
        if best is None or ph["total"] < best["total"]:
            best = ph
    return {k: round(v * 1e3, 3) for k, v in best.items()} | {"pbs": int(info.n_pbs), "levels": int(info.n_levels)}


hay = np.concatenate([text(256), [0]]).astype(np.uint8)
out = {
    "eq_64": run("eq", [np.concatenate([text(64), [0]]).astype(np.uint8)] * 2),
    "contains_256": run("contains", [hay, hay[100:108].copy()]),
}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "api_phases.json"), "w"), indent=1)
