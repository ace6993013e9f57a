"""Latency of ONE dependency level against its size, for both blind-rotation kernels (throughput: one PBS per pair of
warps, four per SM; latency: one PBS per 128-thread CTA, one per SM, key tiles by bulk TMA).  The crossover sets the
engine's default level-size threshold (fhestr_set_br_mode).  usage: python scripts/gpu_latency.py [sizes...]
Writes gpurun_out/level_latency.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhestring_b200.client import ClientKey  # noqa: E402
from fhestring_b200.engine import Engine, single_term_jobs  # noqa: E402

ck = ClientKey(seed=1)
bsk, ksk = ck.server_keys()
SIZES = [int(x) for x in sys.argv[1:]] or [1, 16, 148, 149, 250, 296, 297, 444, 500, 592, 1184]
BMAX = max(1184, max(SIZES))
eng = Engine(arena_blocks=2 * BMAX + 8)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
eng.load_keys(bsk, ksk)
vals = np.random.default_rng(0).integers(0, 16, BMAX).astype(np.uint8)
eng.upload(0, ck.encrypt_blocks(vals))
ident = eng.lut(list(range(16)))
rows = []
for B in SIZES:
    jobs = single_term_jobs(BMAX + np.arange(B), np.arange(B), ident)
    prog = eng.program(jobs, [0, B])
    row = dict(jobs=B)
    for mode, name in ((1, "throughput"), (3, "latency_single"), (4, "latency_pair"), (0, "by_level_size")):
        eng.set_br_mode(mode)
        for _ in range(2):
            prog.run()
        torch.cuda.synchronize()
        eng.set_timing(True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3):
            prog.run()
        b.record(stream)
        torch.cuda.synchronize()
        ks_ms, br_ms, nl, npbs = eng.get_timing()
        eng.set_timing(False)
        ok = bool(np.array_equal(ck.decrypt_blocks(eng.download(BMAX, B)), vals[:B]))
        row[name] = dict(level_ms=round(a.elapsed_time(b) / 3, 4), keyswitch_ms=round(ks_ms / 3, 4),
                         blind_rotate_ms=round(br_ms / 3, 4), decrypt_ok=ok)
    rows.append(row)
    print(json.dumps(row), flush=True)
    prog.close()
eng.set_br_mode(0)
eng.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "level_latency.json"), "w"), indent=1)
