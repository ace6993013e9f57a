"""Latency of one dependency level vs its size, for both blind-rotation kernels (shape 1 = one warp per
polynomial, shape 8 = two warps per polynomial).  usage: python scripts/gpu_latency.py"""
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
BMAX = 4736
eng = Engine(arena_blocks=2 * BMAX + 8)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
eng.load_keys(bsk, ksk)
vals = np.random.default_rng(0).integers(0, 16, BMAX).astype(np.uint8)
eng.upload(0, ck.encrypt_blocks(vals))
ident = eng.lut(list(range(16)))
for B in ([int(x) for x in sys.argv[1:]] or (1, 17, 148, 296, 592, 888, 1184, 2368, 4736)):
    jobs = single_term_jobs(BMAX + np.arange(B), np.arange(B), ident)
    prog = eng.program(jobs, [0, B])
    row = [f"B={B:5d}"]
    for shape in (1,) if len(sys.argv) > 1 else (1, 9):
        eng.set_pbs_per_cta(1)
        eng.set_keyswitch_path(0 if shape == 1 else 1)
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
        ok = np.array_equal(ck.decrypt_blocks(eng.download(BMAX, B)), vals[:B])
        row.append(f"ks {'imma' if shape == 1 else 'imad'}: level {a.elapsed_time(b) / 3:7.3f} ms (ks {ks_ms / 3:6.3f}, br {br_ms / 3:7.3f}) ok={ok}")
    print("  ".join(row), flush=True)
    prog.close()
eng.close()
