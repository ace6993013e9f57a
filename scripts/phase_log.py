"""Timing experiment: when do the PBS that share an SM start their CMUX steps?  Needs a library built with
-DFHESTR_BR_PHASELOG (python fhestring_b200/build.py --variant plog FHESTR_BR_PHASELOG=1 [FHESTR_BR_QUAD=1 ...]) and
FHESTR_ENGINE_LIB pointing at it.  Prints, per SM, the offsets between the logged PBS over the steps."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhestring_b200.client import ClientKey  # noqa: E402
from fhestring_b200.engine import Engine, single_term_jobs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
ck = ClientKey(seed=1)
bsk, ksk = ck.server_keys()
eng = Engine(arena_blocks=2 * B + 8)
eng.load_keys(bsk, ksk)
vals = np.random.default_rng(0).integers(0, 16, B).astype(np.uint8)
eng.upload(0, ck.encrypt_blocks(vals))
ident = eng.lut(list(range(16)))
prog = eng.program(single_term_jobs(B + np.arange(B), np.arange(B), ident), [0, B])
eng.set_br_mode(1)
lib = eng.lib
SMS, STEPS = 8, 1024
MARKS = 18
log = np.zeros((SMS, 4, MARKS, STEPS), np.uint64)
meta = np.zeros((SMS, 4, 4), np.uint32)
for rep in range(2):
    lib.fhestr_debug_phase_log(None, None, 1)
    prog.run()
    torch.cuda.synchronize()
lib.fhestr_debug_phase_log(log.ctypes.data_as(C.c_void_p), meta.ctypes.data_as(C.c_void_p), 0)
cta = np.zeros((8192, 4), np.uint64)
lib.fhestr_debug_phase_log(cta.ctypes.data_as(C.c_void_p), None, 2)
cta = cta[:B].astype(np.int64)
t0 = cta[:, 1].min()
life = (cta[:, 3] - cta[:, 1]) / 1e3
loop_end = (cta[:, 2] - cta[:, 1]) / 1e3
print("per-PBS life (us): min %.0f median %.0f max %.0f; kernel span %.0f us; tail after loop median %.1f us" % (life.min(), np.median(life), life.max(), (cta[:, 3].max() - t0) / 1e3, np.median(life - loop_end)))
logged = cta[:, 0] < 8
print("life on logged SMs: median %.0f us; on the others: median %.0f us" % (np.median(life[logged]), np.median(life[~logged])))
np.save(os.path.join(ROOT, "gpurun_out", f"cta_log_{os.environ.get('PHASE_TAG', 'x')}.npy"), cta)
ok = bool(np.array_equal(ck.decrypt_blocks(eng.download(B, B)), vals))
out = dict(B=B, decrypt_ok=ok, sms=[])
for sm in range(SMS):
    t = log[sm, :, 0, :740].astype(np.int64)      # step starts
    m = log[sm, :, 4, :740].astype(np.int64)      # product-stage starts
    seg = np.diff(np.concatenate([log[sm, :, :10, :700].astype(np.int64), log[sm, :, :1, 1:701].astype(np.int64)], axis=1), axis=1)
    if t[0, 0] == 0:
        continue
    step = np.diff(t, axis=1)
    row = dict(sm=sm, warpid=meta[sm, :, 0].tolist(), cta=meta[sm, :, 1].tolist(), slot=meta[sm, :, 2].tolist(),
               mean_step_cycles=[round(float(x), 1) for x in step.mean(axis=1)],
               fwd_part_cycles=[round(float(x), 1) for x in (m - t).mean(axis=1)],
               # gather | fwd pass 1 | transpose | fwd pass 2 | product | inv pass 1 | transpose | inv pass 2 | accumulate | loop
               segments=[[int(round(float(x))) for x in np.median(seg[k], axis=1)] for k in range(4)])
    prod = np.diff(np.concatenate([log[sm, :, 4:5, :700], log[sm, :, 10:18, :700], log[sm, :, 5:6, :700]], axis=1).astype(np.int64), axis=1)
    # fwd done -> stores | barrier | product 0 | barrier | stores | barrier | product 1 | barrier | -> mark 5
    row["product_parts"] = [[int(round(float(x))) for x in np.median(prod[k], axis=1)] for k in range(4)]
    # offsets of PBS k against PBS 0, in fractions of PBS 0's step, at a few steps
    for s in (1, 10, 50, 100, 200, 400, 700):
        d = (t[:, s] - t[0, s]) / step[0].mean()
        row[f"offset_at_{s}"] = [round(float(x), 3) for x in d]
    out["sms"].append(row)
    print(json.dumps(row))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"phase_log_{os.environ.get('PHASE_TAG', 'x')}.json"), "w"), indent=1)
np.save(os.path.join(ROOT, "gpurun_out", f"phase_log_{os.environ.get('PHASE_TAG', 'x')}.npy"), log[:, :, :, :742])
