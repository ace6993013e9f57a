#!/usr/bin/env python3
"""bench.py -- batched-PBS throughput on B200 (BASELINE.json metric: PBS/s at PARAM_MESSAGE_2_CARRY_2) and the latency
of the encrypted string queries of BASELINE.json configs 3, 4 and 5.

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the CPU path (oracle f64-FFT port) on ALL host cores

Headline workload (BASELINE.json configs[1]): 4096 independent 2-bit-message radix blocks per GPU, identity LUT on
even jobs and the bivariate-eq LUT on odd jobs, PARAM_MESSAGE_2_CARRY_2_KS_PBS (n=742, N=2048).  One step = one pass
of the hot path (keyswitch -> mod-switch -> blind rotation -> sample extract) over the batch.  Independent blocks shard
across ranks with no data-path collective (weak scaling).

String legs (`strings` in the JSON line; levels sharded over the ranks, strong scaling):
  contains_256 / find_256   config 4: encrypted 8-char pattern over a 256-char encrypted string
                            (/root/reference/src/server_key/mod.rs:151-182, :1010-1053)
  eq_64 / ge_64 / le_64     config 3: == / >= / <= on two 64-char strings (mod.rs:1122, :1685, :1613)
  replace_1024              config 5: replace with encrypted from/to over a 1024-char padded string (mod.rs:624-882)
  contains_x16              the throughput form of config 4: 16 queries recorded in one graph
Each query reports three times: `latency_ms` (device-resident inputs, compiled program, CUDA events),
`e2e_latency_ms` (+ H2D of the inputs from pinned memory and D2H of the result) and `api_latency_ms` (the call a
user makes -- MyServerKey.<method>() then the download: recording the op graph, levelising it, binding LUTs,
upload, run, download; host wall clock, max over ranks).

JSON keys beyond the base contract:
  roofline      FP64-FMA roofline of the dominant kernel (blind rotation): algorithmic flops per launch
                (194 510 848 per PBS) / CUDA-event duration of that kernel, against the DFMA
                peak measured on this GPU in the same run (MEASURED_PEAKS.json has no FP64 figure).
  cpu_baseline  the oracle's f64-FFT PBS (a port of the tfhe-rs route; tfhe-rs itself cannot be built
                here) on the host cores, bounded sample.  A reported baseline, not the target.
  e2e           same metric through the C ABI with host buffers: H2D of every batch, PBS, D2H of every result,
                double-buffered over two arena halves and three streams so the copies run under the kernels.
The process exits non-zero when any decrypted result is wrong (the line is still printed, with the flags).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_PBS = 194_510_848          # 742 CMUX x 2^18 FP64 flops
BATCH = 4096                         # BASELINE.json configs[1]
METRIC = "PBS/s (PARAM_MESSAGE_2_CARRY_2_KS_PBS, batched)"
WORKLOAD = ("raw batched PBS microbench: 4096 independent 2-bit-message radix blocks per GPU, "
            "identity/eq LUT, PARAM_MESSAGE_2_CARRY_2 (n=742, N=2048, k=1, PBS 1x23b, KS 5x3b)")
EQ_TABLE = [int((x >> 2) == (x & 3)) for x in range(16)]
# dram__bytes_read.sum + dram__bytes_write.sum of ONE blind_rotate_kernel launch over 4096 PBS, from the committed
# ncu --set full capture (cannot be measured inside a bench run: ncu replays kernels ~40 times)
TRAFFIC_CAPTURE = {"bytes": 180.7e6, "file": "profiles/r2b_blind_rotate_ncu_full.csv (ncu --set full, scripts/r2b_profiles_gpu.sh)",
                   "note": "106.6 MB read + 74.1 MB written; algorithmic 140.0 MB (24.3 MB keyswitched inputs + 67.1 MB "
                           "outputs + 48.6 MB Fourier BSK once)"}


def shared_config(world: int) -> dict:
    """the `config` object, identical for both arms (the reference arm runs the same workload on the host CPU)"""
    return {
        "workload": WORKLOAD, "blocks_per_gpu": BATCH, "luts": ["identity", "eq2"],
        "l2_policy": "inputs larger than L2: 67 MB in + 67 MB out + 49 MB Fourier BSK + 61 MB KSK per step "
                     "(the BSK is meant to be L2-resident inside a launch)",
        "parallelism": f"independent blocks sharded over {world} GPU(s), keys replicated, no collective",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class CpuPort:
    """the oracle's f64-FFT PBS on the host cores (tfhe-rs' own route: keyswitch, mod-switch, fold + twist + FFT
    external products, sample extract; four-step FFT with AVX2 / AVX-512 clones, OpenMP over groups of ciphertexts).
    The thread count is set EXPLICITLY to the cores this process may run on: torchrun exports OMP_NUM_THREADS=1."""

    def __init__(self):
        os.environ["OMP_NUM_THREADS"] = str(host_cores())      # before libgomp initialises
        from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
        self.o = Oracle(**P)
        self.cores = self.o.set_threads(host_cores())
        self.keys = self.o.keygen(1)
        self.fb = self.o.fourier_bsk(self.keys)
        self.luts = np.stack([self.o.lut_poly(list(range(16))), self.o.lut_poly(EQ_TABLE)])
        self.rng = np.random.default_rng(2)

    def run(self, count: int):
        """-> (seconds, threads) for `count` blocks of the bench workload, decrypt-checked"""
        o = self.o
        vals = self.rng.integers(0, 16, count)
        cts = o.encrypt_big(self.keys, vals, seed=9)
        ids = (np.arange(count) % 2).astype(np.int32)
        t = time.perf_counter()
        out, th = o.pbs_fft(self.keys, self.fb, self.luts, ids, cts)
        dt = time.perf_counter() - t
        want = np.where(ids == 1, ((vals >> 2) == (vals & 3)).astype(np.int64), vals)
        if not np.array_equal(o.decrypt_big(self.keys, out), want):
            raise SystemExit("CPU port decrypted wrongly")
        return dt, th


PORT_NOTE = ("oracle f64-FFT PBS: a C port of the tfhe-rs route (four-step FFT, AVX2/AVX-512 clones, OpenMP over groups "
             "of 4 ciphertexts that share every key row); tfhe-rs itself is not buildable here (no Rust toolchain), so "
             "GPU/CPU ratios are against this port")


def cpu_port_rate(seconds_target: float = 12.0):
    """bounded sample of the bench workload -> (PBS/s, cores, sample description)"""
    port = CpuPort()
    dt, th = port.run(4 * port.cores)                                # calibration (also warms caches)
    count = int(max(4 * th, min(BATCH, 4 * th * max(1, round(seconds_target / max(dt, 1e-3))))))
    dt, th = port.run(count)
    return count / dt, th, f"{count} of the {BATCH} blocks, {dt:.1f} s, {PORT_NOTE}"


def reference_arm(args, out):
    """--impl reference: the reference's own CPU implementation is tfhe-rs (Rust, not buildable here: no cargo, crate
    not vendored), so this times the oracle port of the same f64-FFT algorithm on every host core, each step one
    full pass over the same 4096-block batch our arm runs (fewer only if a pass would take more than ~20 s: the
    sample is then stated).  Under torchrun rank 0 alone works; the other ranks exit at once."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    port = CpuPort()
    dt, th = port.run(4 * port.cores)
    per_pass = dt / (4 * port.cores) * BATCH
    blocks = BATCH if per_pass <= 20.0 else int(max(4 * port.cores, BATCH * 20.0 / per_pass) // 4 * 4)
    times = []
    for i in range(args.warmup + args.steps):
        dt, th = port.run(blocks)
        if i >= args.warmup:
            times.append(dt)
    v = blocks * len(times) / sum(times)
    sample = (f"every step: {blocks} of the {BATCH} blocks" + (" (the full batch)" if blocks == BATCH else " (bounded sample)")
              + f", {np.mean(times):.2f} s per step; {PORT_NOTE}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "PBS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(args.gpus),
        "blocks_per_step": blocks,
        "cpu_baseline": {"value": v, "unit": "PBS/s", "cores": th, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=out, flush=True)


# ------------------------------------------------------------------------------------------------ string legs
def _std_expect(method, vals):
    """what Rust std returns for the query (the reference's own self-check, main.rs:47-115)"""
    strs = [bytes(v).split(b"\0")[0].decode("ascii") for v in vals]
    s = strs[0]
    if method == "contains": return int(strs[1] in s)
    if method == "find": return s.find(strs[1]) if strs[1] in s else 255
    if method == "eq": return int(s == strs[1])
    if method == "ge": return int(s >= strs[1])
    if method == "le": return int(s <= strs[1])
    if method == "replace": return s.replace(strs[1], strs[2])
    raise KeyError(method)


def query_leg(ctx, name, method, vals, steps):
    """one encrypted string query, three timings (module docstring).  vals: the arguments as uint8 arrays, strings
    already NUL-padded (STRING_PADDING = 1, main.rs:12), patterns unpadded."""
    import ctypes as C
    import torch
    from fhestring_b200.fhestring import FheAsciiChar, FheString
    from fhestring_b200.graph import Graph
    eng, ck, stream, rank, world, barrier = ctx["eng"], ctx["ck"], ctx["stream"], ctx["rank"], ctx["world"], ctx["barrier"]
    want = _std_expect(method, vals)
    is_str = isinstance(want, str)
    cts_np = ck.encrypt_u8(np.concatenate(vals)).reshape(-1, eng.big)
    cts = torch.from_numpy(cts_np).pin_memory()
    n_in = cts.shape[0]
    in_ptr = C.cast(cts.data_ptr(), C.POINTER(C.c_uint64))

    g = Graph()                                              # one query = one graph (slots restart at 0)
    ids = [g.input_chars(len(v))[0] for v in vals]
    rs, rc = g.string_op(method, ids, fast=True)
    out_ids = list(rs) if is_str else [rc]
    g.mark_output(out_ids)
    info = g.compile(world)
    if info.slots_used > eng.arena_blocks:
        g.close()
        return {"skipped": f"needs {info.slots_used} arena blocks, engine has {eng.arena_blocks}"}
    eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(0), C.c_uint32(n_in), in_ptr))
    eng.sync()
    prog = g.bind(eng)
    res_slots = g.char_slots(out_ids).reshape(-1).astype(np.int64)
    contiguous = bool((np.diff(res_slots) == 1).all())
    slots_u32 = np.ascontiguousarray(res_slots, np.uint32)
    host_res = torch.zeros((len(res_slots), eng.big), dtype=torch.int64).pin_memory()

    def download():
        if contiguous:
            eng._ck(eng.lib.fhestr_ct_download(eng.h, C.c_uint32(int(res_slots[0])), C.c_uint32(len(res_slots)),
                                               C.cast(host_res.data_ptr(), C.POINTER(C.c_uint64))))
        else:
            eng._ck(eng.lib.fhestr_ct_download_slots(eng.h, slots_u32.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(len(res_slots)),
                                                     C.cast(host_res.data_ptr(), C.POINTER(C.c_uint64))))

    reps = max(3, min(steps, 5)) if info.n_pbs < 50_000 else 3
    for _ in range(2 if info.n_pbs < 50_000 else 1):
        prog.run(rank=rank, world=world)
    barrier()
    dev, e2e = [], []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        prog.run(rank=rank, world=world)
        b.record(stream)
        barrier()
        dev.append(a.elapsed_time(b))
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(0), C.c_uint32(n_in), in_ptr))
        prog.run(rank=rank, world=world)
        download()
        b.record(stream)
        barrier()
        e2e.append(a.elapsed_time(b))
    dec = ck.decrypt_u8(host_res.numpy().view(np.uint64).reshape(-1, 4, eng.big))
    got = bytes(dec).split(b"\0")[0].decode("ascii", "replace") if is_str else int(dec[0])
    _, _, npbs, _ = g.program()
    prog.close()
    g.close()

    # the call a user makes: MyServerKey.<method>(...) then the download (record + compile + bind + upload + run + D2H);
    # the chars are views of the PINNED input buffer (the e2e contract's host memory), string by string as they lie
    sk = ctx["sk"]
    cts_pinned = cts.numpy().view(np.uint64)
    api = []
    api_got = None
    for _ in range(2 if info.n_pbs >= 50_000 else 4):
        sk.reset()
        off, args = 0, []
        for i, v in enumerate(vals):
            chars = [FheAsciiChar(ct=cts_pinned[4 * (off + j):4 * (off + j + 1)]) for j in range(len(v))]
            args.append(FheString(chars) if i == 0 or method in ("eq", "ge", "le") else chars)
            off += len(v)
        barrier()
        t0 = time.perf_counter()
        r = getattr(sk, method)(*args)
        outs = list(r.bytes) if is_str else [r]
        raw = sk._download(outs)
        api.append((time.perf_counter() - t0) * 1e3)
        d = ck.decrypt_u8(raw)
        api_got = bytes(d).split(b"\0")[0].decode("ascii", "replace") if is_str else int(d[0])
    barrier()
    # the first call of a query shape records, compiles and binds it; the later ones hit MyServerKey's plan cache
    api_first, api_cached = float(api[0]), float(np.min(api[1:])) if len(api) > 1 else float(api[0])
    t = torch.tensor([float(np.median(dev)), float(np.median(e2e)), api_cached, api_first], device="cuda", dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    show = (lambda x: x if not isinstance(x, str) or len(x) <= 48 else x[:45] + "...")
    return {"latency_ms": float(t[0]), "e2e_latency_ms": float(t[1]), "api_latency_ms": float(t[2]),
            "api_first_call_ms": float(t[3]),
            "api_how": "MyServerKey.<method>(...) + download, host wall clock: api_first_call_ms = the first call of this query "
                       "shape (record + compile + bind + upload + run + D2H), api_latency_ms = the best later call on the "
                       "same shape (plan cache: upload + run of the bound program + D2H)",
            "levels": int(info.n_levels), "pbs": int(info.n_pbs), "level_pbs": [int(x) for x in npbs],
            "decrypted": show(got), "expected": show(want), "verified": bool(got == want and api_got == want),
            "h2d_bytes": int(n_in * eng.big * 8), "d2h_bytes": int(len(res_slots) * eng.big * 8),
            "reference": f"/root/reference/src/server_key/mod.rs ({method})"}


def strings_leg(ctx, steps, cpu_rate):
    import torch
    from fhestring_b200.graph import Graph
    eng, ck, stream, rank, world, barrier = ctx["eng"], ctx["ck"], ctx["stream"], ctx["rank"], ctx["world"], ctx["barrier"]
    out = {}
    rng = np.random.default_rng(4)
    alpha = list(b"abcdefghilmnoprstu")
    pad = np.zeros(1, np.uint8)
    # config 4
    body = rng.choice(alpha, 256).astype(np.uint8)
    pat = np.frombuffer(b"qzjxkvwq", np.uint8)
    body[124:132] = pat
    s256 = np.concatenate([body, pad])
    out["contains_256"] = query_leg(ctx, "contains_256", "contains", [s256, pat], steps)
    out["find_256"] = query_leg(ctx, "find_256", "find", [s256, pat], steps)
    # config 3: two 64-char strings that share a 40-char prefix
    a64 = rng.choice(alpha, 64).astype(np.uint8)
    b64 = a64.copy()
    b64[40:] = rng.choice(alpha, 24)
    b64[40] = a64[40] + 1
    for m in ("eq", "ge", "le"):
        out[f"{m}_64"] = query_leg(ctx, f"{m}_64", m, [np.concatenate([a64, pad]), np.concatenate([b64, pad])], steps)
    # config 5: replace over 1024 chars, 8 plants of a 4-char pattern
    body = rng.choice(alpha, 1024).astype(np.uint8)
    frm, to = np.frombuffer(b"qzjx", np.uint8), np.frombuffer(b"WXYZ", np.uint8)
    for k in range(8):
        body[100 + 120 * k:104 + 120 * k] = frm
    out["replace_1024"] = query_leg(ctx, "replace_1024", "replace", [np.concatenate([body, pad]), frm, to], steps)
    # throughput form of config 4: Q independent queries recorded in ONE graph -- their levels merge, so every level is
    # Q times wider and fills all ranks
    Q = 16
    g = Graph()
    outs, bodies = [], []
    rngq = np.random.default_rng(44)
    for qi in range(Q):
        b = rngq.choice(alpha, 256).astype(np.uint8)
        if qi % 2 == 0:
            b[(7 * qi) % 248:(7 * qi) % 248 + 8] = pat
        bodies.append(np.concatenate([b, pad]))
    ids_p, slots_p = g.input_chars(len(pat))
    first_slot = int(slots_p[0, 0])
    all_vals = [pat]
    for qi in range(Q):
        ids_s, _ = g.input_chars(len(bodies[qi]))
        all_vals.append(bodies[qi])
        _, cid = g.string_op("contains", [ids_s, ids_p], fast=True)
        outs.append(cid)
    g.mark_output(outs)
    info = g.compile(world)
    if info.slots_used <= eng.arena_blocks:
        eng.upload(first_slot, ck.encrypt_u8(np.concatenate(all_vals)).reshape(-1, eng.big))
        prog = g.bind(eng)
        for _ in range(2):
            prog.run(rank=rank, world=world)
        barrier()
        times = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            prog.run(rank=rank, world=world)
            b.record(stream)
            barrier()
            times.append(a.elapsed_time(b))
        t = torch.tensor([float(np.median(times))], device="cuda", dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res_slots = g.char_slots(outs)[:, 0]
        got = [int(ck.decrypt_blocks(eng.download(int(sl), 1))[0]) for sl in res_slots]
        want = [1 if qi % 2 == 0 else 0 for qi in range(Q)]
        _, _, npbs, _ = g.program()
        out["contains_x16"] = {"queries": Q, "ms": float(t[0]), "queries_per_s": Q / (float(t[0]) * 1e-3),
                               "pbs": int(info.n_pbs), "levels": int(info.n_levels), "level_pbs": [int(x) for x in npbs],
                               "pbs_per_s": int(info.n_pbs) / (float(t[0]) * 1e-3), "verified": bool(got == want)}
        prog.close()
    g.close()
    # the CPU side of config 4: the reference issues its own op order (12 250 non-trivial PBS in 260
    # dependent levels for contains over 257 x 8 chars); at the port's measured all-core PBS rate that is
    if cpu_rate:
        out["cpu_estimate"] = {
            "contains_256_reference_graph_s": 12250 / cpu_rate, "contains_256_this_graph_s": out["contains_256"].get("pbs", 0) / cpu_rate,
            "how": "PBS count of the graph / the CPU port's measured PBS rate on all host cores in this run (cpu_baseline); "
                   "an extrapolation, not a timed run -- the reference's 260 dependent levels would also cap its parallelism",
            "reference_graph": {"pbs_nontrivial": 12250, "pbs_nominal": 19000, "levels": 260, "source": "SURVEY.md 2.6"}}
    out["workload"] = ("configs 3, 4, 5 of BASELINE.json as depth-minimised graphs, levels sharded over the ranks; small levels "
                       "run on the latency kernel (one PBS per SM), large ones on the throughput kernel")
    return out


def _tensor_side(pbs_per_launch, ks_ms_per_launch):
    """keyswitch (ks_digits_kernel + ks_gemm_tc_kernel) per launch against the tensor peak: MEASURED_PEAKS.json's dense bf16
    figure x 2 (the int8 rate of B200's tensor cores is twice the bf16 rate), else nominal 4 500 TOP/s"""
    if ks_ms_per_launch <= 0:
        return None
    ops = 2.0 * pbs_per_launch * (742 + 1) * 8 * 2048 * 5          # M x (n + 1) limb columns x K, multiply + add
    peak, src = 4500.0, "nominal dense int8"
    try:
        peak = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        src = "2 x MEASURED_PEAKS.json bf16_tflops"
    except Exception:
        pass
    achieved = ops / (ks_ms_per_launch * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "ks_gemm_tc_kernel (+ ks_digits_kernel)", "achieved": achieved, "peak": peak,
            "unit": "TOP/s", "frac": achieved / peak, "peak_source": src}


def _hbm_side(traffic_bytes, ms_per_launch):
    """DRAM traffic of one blind-rotation launch (ncu) over its duration, against MEASURED_PEAKS.json's hbm_gbs"""
    if not traffic_bytes or ms_per_launch <= 0:
        return None
    peak, src = 7700.0, "nominal HBM3e"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    achieved = traffic_bytes / (ms_per_launch * 1e-3) / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": src}


def _claim_stdout():
    """the contract is ONE JSON line on stdout: keep a private handle to it and point fd 1 at stderr so that
    library banners (NCCL prints its version on stdout) cannot pollute it"""
    sys.stdout.flush()
    keep = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return keep


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strings", "--no-contains", dest="no_strings", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU level exchange: P2P stores from the kernel epilogue + flag barrier (default) or NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args, out)
    args.warmup = max(args.warmup, 3)

    import ctypes as C
    import torch
    import torch.distributed as dist
    from fhestring_b200.client import ClientKey
    from fhestring_b200.engine import Engine, single_term_jobs
    from fhestring_b200.fhestring import MyServerKey

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch

    # ---- setup (untimed): keys, ciphertexts, engine.  The arena also holds config 5 (replace over 1025 chars needs
    # about 216 k blocks = 3.5 GB)
    ck = ClientKey(seed=1)          # benchmark keys: an explicit test seed (the default is OS entropy)
    bsk, ksk = ck.server_keys()
    eng = Engine(arena_blocks=max(4 * B + 8, 1 << 18), device=local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.load_keys(bsk, ksk)
    rng = np.random.default_rng(1000 + rank)
    vals = rng.integers(0, 16, B).astype(np.uint8)
    cts = ck.encrypt_blocks(vals)
    ident, eq = eng.lut(list(range(16))), eng.lut(EQ_TABLE)
    want = np.where(np.arange(B) % 2 == 1, ((vals >> 2) == (vals & 3)).astype(np.uint8), vals)
    # two arena halves: inputs [0, B) -> results [2B, 3B), inputs [B, 2B) -> results [3B, 4B)
    progs = []
    for h in range(2):
        jobs = single_term_jobs(2 * B + h * B + np.arange(B), h * B + np.arange(B), ident)
        jobs["lut"][1::2] = eq
        progs.append(eng.program(jobs, [0, B]))
    eng.upload(0, cts)
    prog = progs[0]
    fp64_peak, _ = eng.measure_fp64_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: inputs already in HBM
    for _ in range(args.warmup):
        prog.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.set_timing(True)
    launches0 = eng.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        prog.run()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ks_ms, br_ms, br_launches, br_pbs = eng.get_timing()
    eng.set_timing(False)
    gpu_launches = eng.kernel_launches() - launches0
    got = ck.decrypt_blocks(eng.download(2 * B, B))
    verified = bool(np.array_equal(got, want))

    # ---- end-to-end leg: host buffers through the C ABI; every step uploads its batch from pinned memory and
    # downloads its results.  Double-buffered: step i uses arena half i & 1; uploads, kernels and downloads run on
    # three streams ordered by events, so the copies of steps i+1 and i-1 run under the kernels of step i.
    host_in = torch.from_numpy(cts).pin_memory()
    host_out = [torch.empty_like(host_in).pin_memory() for _ in range(2)]
    in_ptr = C.cast(host_in.data_ptr(), C.POINTER(C.c_uint64))
    out_ptr = [C.cast(t.data_ptr(), C.POINTER(C.c_uint64)) for t in host_out]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def e2e_pipeline(n_steps):
        ev_in = [None, None]; ev_comp = [None, None]; ev_out = [None, None]
        for i in range(n_steps):
            h = i & 1
            if ev_comp[h] is not None:
                s_in.wait_event(ev_comp[h])                       # the kernels that read this input half are done
            eng.set_stream(s_in.cuda_stream)
            eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(h * B), C.c_uint32(B), in_ptr))
            ev_in[h] = torch.cuda.Event(); ev_in[h].record(s_in)
            stream.wait_event(ev_in[h])
            if ev_out[h] is not None:
                stream.wait_event(ev_out[h])                      # the previous results of this half have left
            eng.set_stream(stream.cuda_stream)
            progs[h].run()
            ev_comp[h] = torch.cuda.Event(); ev_comp[h].record(stream)
            s_out.wait_event(ev_comp[h])
            eng.set_stream(s_out.cuda_stream)
            eng._ck(eng.lib.fhestr_ct_download_async(eng.h, C.c_uint32(2 * B + h * B), C.c_uint32(B), out_ptr[h]))
            ev_out[h] = torch.cuda.Event(); ev_out[h].record(s_out)
        eng.set_stream(stream.cuda_stream)
        for h in range(2):
            if ev_out[h] is not None:
                stream.wait_event(ev_out[h])

    e2e_pipeline(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    e2e_pipeline(args.steps)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_verified = all(bool(np.array_equal(ck.decrypt_blocks(t.numpy()), want)) for t in host_out[:min(2, args.steps)])
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline (rank 0, bounded sample) -- before the string legs, which quote it
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, cores, sample = cpu_port_rate(12.0)
        cpu = {"value": v, "unit": "PBS/s", "cores": cores, "kind": "port", "sample": sample}

    # ---- string legs: configs 3, 4, 5, levels sharded over the ranks
    strings = None
    if not args.no_strings:
        exchange = "none (1 GPU)"
        if world > 1:
            if args.exchange == "p2p":
                eng.peer_attach(rank, world)
                exchange = "blind-rotation epilogue stores results into every peer arena over NVLink (cudaIpc) + flag barrier per level"
            else:
                eng.comm_init(rank, world)
                exchange = "in-place ncclAllGather per level"
        sk = MyServerKey(None, None, engine=eng, rank=rank, world=world)
        ctx = dict(eng=eng, ck=ck, stream=stream, rank=rank, world=world, barrier=barrier, sk=sk)
        strings = strings_leg(ctx, args.steps, cpu["value"] if cpu else None)
        strings["exchange"] = exchange
        if world > 1 and args.exchange == "p2p" and eng.peer_timed_out():
            strings["peer_barrier_timed_out"] = True
    del bsk, ksk

    # ---- max over ranks
    t = torch.tensor([ms_total, e2e_ms, br_ms], device="cuda", dtype=torch.float64)
    strings_ok = strings is None or (all(v.get("verified", True) for v in strings.values() if isinstance(v, dict))
                                     and not strings.get("peer_barrier_timed_out"))
    ok = torch.tensor([int(verified), int(e2e_verified), int(strings_ok)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms_total, e2e_ms, br_ms_max = [float(x) for x in t.tolist()]
    verified, e2e_verified, strings_ok = [bool(x) for x in ok.tolist()]

    if rank == 0:
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
        br_avg_ms = br_ms / max(1, br_launches)
        achieved = (br_pbs / max(1, br_launches)) * FLOPS_PER_PBS / (br_avg_ms * 1e-3) / 1e12
        traffic = TRAFFIC_CAPTURE["bytes"] if B == BATCH else None
        line = {
            "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(world),
            "verified_decrypt": verified, "e2e_verified": e2e_verified, "strings_verified": strings_ok,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                "traffic": traffic, "traffic_unit": "bytes per launch",
                "traffic_source": {k: v for k, v in TRAFFIC_CAPTURE.items() if k != "bytes"},
                "kernel": "blind_rotate_kernel", "ms_per_launch": br_avg_ms,
                "flops_per_launch": (br_pbs / max(1, br_launches)) * FLOPS_PER_PBS,
                "peak_source": "DFMA microbenchmark measured in this run (fhestr_measure_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure; nominal 37.2 TFLOP/s",
                "keyswitch_ms_per_launch": ks_ms / max(1, br_launches),
                # the second kernel pair of a step against ITS roofline: the limb-split GEMM on tcgen05.mma kind::i8
                # (u8 MACs of the real 743 x 8 limb columns; the time includes the digits kernel in front of the GEMM)
                "keyswitch": _tensor_side(br_pbs / max(1, br_launches), ks_ms / max(1, br_launches)),
                "kernel_share_of_step": br_ms / ms_total,
                # why the bound is FP64 and not HBM: the same kernel against the measured copy bandwidth
                "hbm": _hbm_side(traffic, br_avg_ms),
            },
            "e2e": {"value": e2e_value, "unit": "PBS/s", "h2d_bytes_per_step": int(B * 2049 * 8),
                    "d2h_bytes_per_step": int(B * 2049 * 8), "ms_per_step": e2e_ms / args.steps,
                    "how": "every step: H2D of the batch from pinned memory, PBS, D2H of the results; two arena halves, "
                           "three streams (copies of steps i+1 / i-1 under the kernels of step i)"},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
        }
        if strings is not None:
            line["strings"] = strings
            line["contains_256"] = strings.get("contains_256")    # round-1 key, kept for comparison across rounds
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    if not (verified and e2e_verified and strings_ok):
        sys.exit(3)


if __name__ == "__main__":
    main()
