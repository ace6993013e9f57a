"""Ad-hoc GPU bring-up check (test infrastructure, lives under tests/ because it uses the oracle as the checker):
parity of each kernel vs the oracle + timings of the blind-rotation launch shapes.
usage: python tests/tools/gpu_check.py [batch] [--skip-small]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from fhestring_b200.engine import Engine, make_jobs, single_term_jobs, PARAM_MESSAGE_2_CARRY_2_KS_PBS as PE
from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as PO

def log2(x): return float(np.log2(np.maximum(np.abs(x).astype(float), 1e-300)))

def small_checks():
    n = 8
    po = dict(PO); po.update(n=n)
    pe = dict(PE); pe.update(n=n)
    o = Oracle(**po)
    keys = o.keygen(11)
    eng = Engine(arena_blocks=256, **pe)
    eng.load_keys(keys.bsk, keys.ksk)
    # LUT bit-exact
    table = [(3 * x + 1) % 16 for x in range(16)]
    lid = eng.lut(table)
    assert np.array_equal(eng.lut_download(lid), o.lut_poly(table)), "LUT poly mismatch"
    print("K5 lut poly: bit-exact")
    vals = np.arange(32) % 16
    cts = o.encrypt_big(keys, vals, seed=5)
    eng.upload(0, cts)
    assert np.array_equal(eng.download(0, 32), cts)
    # keyswitch bit-exact incl. linear combination
    jobs = make_jobs(32)
    for i in range(32):
        jobs[i]["dst"] = 64 + i; jobs[i]["lut"] = lid; jobs[i]["n_terms"] = 2
        jobs[i]["src"][0] = i; jobs[i]["src"][1] = (i + 1) % 32
        jobs[i]["coeff"][0] = 4; jobs[i]["coeff"][1] = -1 if i % 2 else 1
        jobs[i]["constant"] = np.uint64((i % 3) << 59)
    with np.errstate(over="ignore"):
        lin = np.stack([cts[i] * np.uint64(4) + (cts[(i + 1) % 32] * np.uint64(2**64 - 1 if i % 2 else 1)) for i in range(32)])
        for i in range(32):
            lin[i, -1] += np.uint64((i % 3) << 59)
    ks_want = o.keyswitch(keys, lin)
    ks_got = eng.debug_keyswitch(jobs)
    assert np.array_equal(ks_got, ks_want), "keyswitch mismatch"
    print("K0+K1 keyswitch: bit-exact")
    # single CMUX tolerance
    rng = np.random.default_rng(3)
    glwe = rng.integers(0, 2**64, size=(4, 2, 2048), dtype=np.uint64)
    one = dict(pe); one.update(n=1)
    eng1 = Engine(arena_blocks=8, **one)
    eng1.load_keys(keys.bsk[:1], keys.ksk[:, :, [0, n]].copy())
    ks = np.zeros((4, 2), np.uint64)
    es = [1, 777, 2048 + 5, 4095]
    for b, e in enumerate(es): ks[b, 0] = np.uint64(e) << np.uint64(52)
    got = eng1.debug_blind_rotate(ks, None, glwe)
    def monomial(poly, e, N=2048):
        j = np.arange(N); q = (j - e) % (2 * N); v = poly[q % N]
        with np.errstate(over="ignore"):
            return np.where(q >= N, np.uint64(0) - v, v)
    for b, e in enumerate(es):
        with np.errstate(over="ignore"):
            diff = np.stack([monomial(glwe[b, r], e) - glwe[b, r] for r in range(2)])
        want = o.external_product_exact(keys.bsk[0], diff, glwe[b])
        d = (got[b] - want).astype(np.int64)
        print(f"K3 single CMUX e={e}: max 2^{log2(np.abs(d).max()):.2f} rms 2^{log2(np.sqrt(np.mean(d.astype(float)**2))):.2f} (units of 2^-64 torus)")
    eng1.close()
    # full PBS small n
    pj = single_term_jobs(100 + np.arange(32), np.arange(32), lid)
    eng.pbs_batch(pj)
    out = eng.download(100, 32)
    dec = o.decrypt_big(keys, out)
    want = np.array([table[v] for v in vals])
    print("PBS small n decrypt ok:", np.array_equal(dec, want), dec[:16])
    eng.close()

def full_checks(B=4096):
    o = Oracle(**PO)
    t = time.time(); keys = o.keygen(1); print("keygen", round(time.time() - t, 2), "s")
    eng = Engine(arena_blocks=2 * B + 16, **PE)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    t = time.time(); eng.load_keys(keys.bsk, keys.ksk); print("load_keys", round(time.time() - t, 2), "s")
    rng = np.random.default_rng(2)
    vals = rng.integers(0, 16, B)
    t = time.time(); cts = o.encrypt_big(keys, vals, seed=9); print("encrypt", round(time.time() - t, 2), "s")
    eng.upload(0, cts)
    ident = eng.lut(list(range(16)))
    eq = eng.lut([int((x >> 2) == (x & 3)) for x in range(16)])
    jobs = single_term_jobs(B + np.arange(B), np.arange(B), ident)
    jobs["lut"][1::2] = eq
    tf, mhz = eng.measure_fp64_peak(); print("measured FP64 peak TFLOP/s", tf, "clock attr MHz", mhz)
    for P in (1, 2):
        eng.set_br_mode(P)
        prog = eng.program(jobs, [0, B])
        for _ in range(2): prog.run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); prog.run(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"P={P}: {B} PBS in {ms:.2f} ms -> {B / ms * 1e3:.0f} PBS/s ; FP64 {B / ms * 1e3 * 194510848 / 1e12:.2f} TFLOP/s")
    out = eng.download(B, B)
    dec = o.decrypt_big(keys, out)
    want = np.where(np.arange(B) % 2 == 1, ((vals >> 2) == (vals & 3)).astype(np.int64), vals)
    print("full PBS decrypt ok:", np.array_equal(dec, want), "mismatches", int((dec != want).sum()))
    ph = o.phases(keys.s_glwe, out).astype(np.int64) - (want.astype(np.int64) << 59)
    print("output noise std 2^%.2f of torus" % (log2(np.std(ph.astype(float))) - 64))
    # KS-only timing
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if "--skip-small" not in sys.argv: small_checks()
    full_checks(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
