"""Per-level device time of contains()/find() on a 256-char encrypted string with an encrypted 8-char pattern
(BASELINE config 4): each dependency level is run on its own between CUDA events, so the numbers show where the
latency of one query goes.  usage: python scripts/contains_levels.py [--json out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhestring_b200.client import ClientKey  # noqa: E402
from fhestring_b200.engine import Engine  # noqa: E402
from fhestring_b200.graph import Graph  # noqa: E402

ck = ClientKey(seed=1)
bsk, ksk = ck.server_keys()
eng = Engine(arena_blocks=1 << 15)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
eng.load_keys(bsk, ksk)
rng = np.random.default_rng(4)
body = rng.choice(list(b"abcdefghilmnoprstu"), 256).astype(np.uint8)
pat = np.frombuffer(b"qzjxkvwq", np.uint8)
body[124:132] = pat
s = np.concatenate([body, np.zeros(1, np.uint8)])
cts = ck.encrypt_u8(np.concatenate([s, pat])).reshape(-1, eng.big)
out = {}
for name, want in (("contains", 1), ("find", 124)):
    g = Graph()
    ids_s, _ = g.input_chars(len(s))
    ids_p, _ = g.input_chars(len(pat))
    _, cid = g.string_op(name, [ids_s, ids_p], fast=True)
    g.mark_output([cid])
    info = g.compile(1)
    eng.upload(0, cts)
    prog = g.bind(eng)
    _, _, npbs, _ = g.program()
    for _ in range(2):
        prog.run()
    torch.cuda.synchronize()
    rows = []
    for l in range(info.n_levels):
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            prog.run(l, l + 1)
            b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        rows.append({"level": l, "pbs": int(npbs[l]), "ms": float(np.median(ts))})
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); prog.run(); b.record(stream); torch.cuda.synchronize()
    slot = int(g.char_slots([cid])[0, 0])
    got = int(ck.decrypt_blocks(eng.download(slot, 1))[0]) if name == "contains" else None
    out[name] = {"levels": rows, "whole_program_ms": a.elapsed_time(b), "sum_of_levels_ms": sum(r["ms"] for r in rows)}
    print(name, "whole", round(out[name]["whole_program_ms"], 2), "ms;", "  ".join(f"L{r['level']}: {r['pbs']} PBS {r['ms']:.2f} ms" for r in rows), "decrypted" if got is None else f"decrypted {got} (want {want})")
    prog.close(); g.close()
if "--json" in sys.argv:
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
        json.dump(out, f, indent=1)
eng.close()
