#!/usr/bin/env python3
"""bench.py -- batched-PBS throughput on B200 (BASELINE.json metric: PBS/s at PARAM_MESSAGE_2_CARRY_2).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the CPU path (oracle f64-FFT port) on host cores

Workload (BASELINE.json configs[1]): 4096 independent 2-bit-message radix blocks per GPU, identity LUT on
even jobs and the bivariate-eq LUT on odd jobs, PARAM_MESSAGE_2_CARRY_2_KS_PBS (n=742, N=2048).  One step =
one pass of the hot path (keyswitch -> mod-switch -> blind rotation -> sample extract) over the batch.
A second leg (contains_256) measures the other half of the BASELINE metric: contains()/find() latency on a
256-char encrypted string with an encrypted 8-char pattern (config 4), levels sharded over the ranks, plus the
same workload batched 16 queries wide.
Independent blocks shard across ranks with no data-path collective (weak scaling).

JSON keys beyond the base contract:
  roofline      FP64-FMA roofline of the dominant kernel (blind rotation): algorithmic flops per launch
                (194 510 848 per PBS) / CUDA-event duration of that kernel, against the DFMA
                peak measured on this GPU in the same run (MEASURED_PEAKS.json has no FP64 figure).
  cpu_baseline  the oracle's f64-FFT PBS (a port of the tfhe-rs route; tfhe-rs itself cannot be built
                here) on the host cores, bounded sample.  A reported baseline, not the target.
  e2e           same metric through the C ABI with host buffers: H2D of the batch, PBS, D2H of the results.
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


def cpu_port_rate(seconds_target: float = 12.0):
    """oracle f64-FFT PBS on the host cores, bounded sample of the same workload -> (PBS/s, cores, sample)"""
    from oracle.tfhe_oracle import Oracle, PARAM_MESSAGE_2_CARRY_2_KS_PBS as P
    o = Oracle(**P)
    keys = o.keygen(1)
    fb = o.fourier_bsk(keys)
    threads = o.max_threads()
    luts = np.stack([o.lut_poly(list(range(16))), o.lut_poly(EQ_TABLE)])
    rng = np.random.default_rng(2)

    def run(count):
        vals = rng.integers(0, 16, count)
        cts = o.encrypt_big(keys, vals, seed=9)
        ids = (np.arange(count) % 2).astype(np.int32)
        t = time.perf_counter()
        out, th = o.pbs_fft(keys, fb, luts, ids, cts)
        dt = time.perf_counter() - t
        want = np.where(ids == 1, ((vals >> 2) == (vals & 3)).astype(np.int64), vals)
        assert np.array_equal(o.decrypt_big(keys, out), want), "CPU port decrypted wrongly"
        return dt, th

    dt, th = run(threads)                      # calibration (also warms caches)
    count = int(max(threads, min(BATCH, threads * max(1, round(seconds_target / max(dt, 1e-3))))))
    dt, th = run(count)
    return count / dt, th, (f"{count} of the {BATCH} blocks, {dt:.1f} s, oracle f64-FFT PBS (OpenMP over ciphertexts); a plain "
                            "radix-2 C port of the tfhe-rs route -- tfhe-rs' own concrete-fft is several times faster per core "
                            "(not measurable here: no Rust toolchain), so treat GPU/CPU ratios as upper bounds")


def contains_leg(eng, ck, stream, rank, world, steps, barrier):
    """BASELINE metric, second half: contains() latency on a 256-char encrypted string with an encrypted 8-char
    pattern (config 4; /root/reference/src/server_key/mod.rs:151-182), recorded depth-minimised, every level's
    PBS jobs sharded over the ranks with one in-place NCCL all-gather per level (strong scaling)."""
    import ctypes as C
    import torch
    from fhestring_b200.graph import Graph
    rng = np.random.default_rng(4)
    body = rng.choice(list(b"abcdefghilmnoprstu"), 256).astype(np.uint8)
    pat = np.frombuffer(b"qzjxkvwq", np.uint8)
    body[124:132] = pat
    s = np.concatenate([body, np.zeros(1, np.uint8)])           # STRING_PADDING = 1 (main.rs:12)
    out = {}
    cts = torch.from_numpy(ck.encrypt_u8(np.concatenate([s, pat])).reshape(-1, eng.big)).pin_memory()
    n_in = cts.shape[0]
    in_ptr = C.cast(cts.data_ptr(), C.POINTER(C.c_uint64))
    eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(0), C.c_uint32(n_in), in_ptr))
    eng.sync()
    for name, want in (("contains", 1), ("find", 124)):
        g = Graph()                                              # one query = one graph (slots restart at 0)
        ids_s, slots_s = g.input_chars(len(s))
        ids_p, slots_p = g.input_chars(len(pat))
        assert int(slots_s[0, 0]) == 0 and int(slots_p[-1, -1]) == n_in - 1
        _, cid = g.string_op(name, [ids_s, ids_p], fast=True)
        g.mark_output([cid])
        info = g.compile(world)
        prog = g.bind(eng)
        res_slots = [int(x) for x in g.char_slots([cid])[0]]
        host_res = torch.zeros((4, eng.big), dtype=torch.int64).pin_memory()
        res_ptrs = [C.cast(host_res[blk].data_ptr(), C.POINTER(C.c_uint64)) for blk in range(4)]
        for _ in range(2):
            prog.run(rank=rank, world=world)
        barrier()
        dev, e2e = [], []
        for _ in range(max(3, steps)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            prog.run(rank=rank, world=world)
            b.record(stream)
            barrier()
            dev.append(a.elapsed_time(b))
        for _ in range(max(3, steps)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(0), C.c_uint32(n_in), in_ptr))
            prog.run(rank=rank, world=world)
            for blk in range(4):   # the result char: 4 radix blocks, each its own arena slot
                eng._ck(eng.lib.fhestr_ct_download(eng.h, C.c_uint32(res_slots[blk]), C.c_uint32(1), res_ptrs[blk]))
            b.record(stream)
            barrier()
            e2e.append(a.elapsed_time(b))
        got = int(ck.decrypt_u8(host_res.numpy().view(np.uint64).reshape(1, 4, eng.big))[0])
        t = torch.tensor([float(np.median(dev)), float(np.median(e2e))], device="cuda", dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        jobs, off, npbs, _ = g.program()
        out[name] = {"latency_ms": float(t[0]), "e2e_latency_ms": float(t[1]), "levels": int(info.n_levels),
                     "pbs": int(info.n_pbs), "level_pbs": [int(x) for x in npbs], "decrypted": got, "expected": want,
                     "verified": bool(got == want), "h2d_bytes": int(n_in * eng.big * 8), "d2h_bytes": int(4 * eng.big * 8)}
        prog.close()
        g.close()
    # throughput form of the same workload: Q independent queries recorded in ONE graph -- their levels merge, so
    # every level is Q times wider and fills all ranks (the single query above is bounded by 6 PBS latencies)
    Q = 16
    g = Graph()
    outs = []
    rngq = np.random.default_rng(44)
    bodies = []
    for qi in range(Q):
        b = rngq.choice(list(b"abcdefghilmnoprstu"), 256).astype(np.uint8)
        if qi % 2 == 0:
            b[(7 * qi) % 248:(7 * qi) % 248 + 8] = pat
        bodies.append(np.concatenate([b, np.zeros(1, np.uint8)]))
    ids_p, slots_p = g.input_chars(len(pat))
    first_slot = int(slots_p[0, 0])
    all_vals = [pat]
    for qi in range(Q):
        ids_s, slots_s = g.input_chars(len(bodies[qi]))
        all_vals.append(bodies[qi])
        _, cid = g.string_op("contains", [ids_s, ids_p], fast=True)
        outs.append(cid)
    g.mark_output(outs)
    info = g.compile(world)
    if info.slots_used <= eng.arena_blocks:
        ctsq = ck.encrypt_u8(np.concatenate(all_vals)).reshape(-1, eng.big)
        eng.upload(first_slot, ctsq)
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
    out["workload"] = ("contains/find, encrypted 8-char pattern over a 256-char encrypted string (+1 NUL padding), "
                       "depth-minimised graph, levels sharded over the ranks")
    out["reference_graph"] = {"contains_pbs_nominal": 19000, "contains_levels": 260, "source": "SURVEY.md 2.6"}
    return out


def _hbm_side(traffic_bytes, ms_per_launch):
    """DRAM traffic of one blind-rotation launch (ncu) over its duration, against MEASURED_PEAKS.json's hbm_gbs"""
    if not traffic_bytes or ms_per_launch <= 0:
        return None
    peak, src = 7700.0, "nominal HBM3e"
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    achieved = traffic_bytes / (ms_per_launch * 1e-3) / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": src}


def reference_arm(args, out):
    """--impl reference: the reference's own CPU implementation is tfhe-rs (Rust, not buildable here: no
    cargo, crate not vendored), so this times the oracle port of the same f64-FFT algorithm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    cores, sample = 1, ""
    per_step = max(4.0, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        r, cores, sample = cpu_port_rate(per_step)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "PBS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "each step is a bounded sample of the 4096-block batch on the host CPU"},
        "cpu_baseline": {"value": v, "unit": "PBS/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=out, flush=True)


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
    ap.add_argument("--no-contains", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU level exchange: P2P stores from the kernel epilogue + flag barrier (default) or NCCL all-gather")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args, out)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from fhestring_b200.client import ClientKey
    from fhestring_b200.engine import Engine, single_term_jobs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch

    # ---- setup (untimed): keys, ciphertexts, engine
    ck = ClientKey(seed=1)
    bsk, ksk = ck.server_keys()
    eng = Engine(arena_blocks=max(2 * B + 8, 1 << 17), device=local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.load_keys(bsk, ksk)
    del bsk, ksk
    rng = np.random.default_rng(1000 + rank)
    vals = rng.integers(0, 16, B).astype(np.uint8)
    cts = ck.encrypt_blocks(vals)
    ident, eq = eng.lut(list(range(16))), eng.lut(EQ_TABLE)
    jobs = single_term_jobs(B + np.arange(B), np.arange(B), ident)
    jobs["lut"][1::2] = eq
    want = np.where(np.arange(B) % 2 == 1, ((vals >> 2) == (vals & 3)).astype(np.uint8), vals)
    eng.upload(0, cts)
    prog = eng.program(jobs, [0, B])
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
    got = ck.decrypt_blocks(eng.download(B, B))
    verified = bool(np.array_equal(got, want))

    # ---- end-to-end leg: host buffers through the C ABI, H2D + PBS + D2H inside the timed region
    host_in = torch.from_numpy(cts).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    import ctypes as C
    in_ptr = C.cast(host_in.data_ptr(), C.POINTER(C.c_uint64))
    out_ptr = C.cast(host_out.data_ptr(), C.POINTER(C.c_uint64))

    def e2e_step():
        eng._ck(eng.lib.fhestr_ct_upload(eng.h, C.c_uint32(0), C.c_uint32(B), in_ptr))
        prog.run()
        eng._ck(eng.lib.fhestr_ct_download(eng.h, C.c_uint32(B), C.c_uint32(B), out_ptr))  # synchronises

    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_verified = bool(np.array_equal(ck.decrypt_blocks(host_out.numpy()), want))
    clocks = sampler.stop() if rank == 0 else None

    # ---- string leg: contains()/find() latency, levels sharded over the ranks
    contains = None
    if not args.no_contains:
        exchange = "none (1 GPU)"
        if world > 1:
            if args.exchange == "p2p":
                eng.peer_attach(rank, world)
                exchange = "blind-rotation epilogue stores results into every peer arena over NVLink (cudaIpc) + flag barrier per level"
            else:
                eng.comm_init(rank, world)
                exchange = "in-place ncclAllGather per level"
        contains = contains_leg(eng, ck, stream, rank, world, args.steps, barrier)
        contains["exchange"] = exchange
        if world > 1 and args.exchange == "p2p":
            assert not eng.peer_timed_out(), "a peer barrier timed out"

    # ---- max over ranks
    t = torch.tensor([ms_total, e2e_ms, br_ms], device="cuda", dtype=torch.float64)
    ok = torch.tensor([int(verified and e2e_verified)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    ms_total, e2e_ms, br_ms_max = [float(x) for x in t.tolist()]
    all_ok = bool(ok.item())

    if rank == 0:
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
        br_avg_ms = br_ms / max(1, br_launches)
        achieved = (br_pbs / max(1, br_launches)) * FLOPS_PER_PBS / (br_avg_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "blocks_per_gpu": B, "luts": ["identity", "eq2"],
                "l2_policy": "inputs larger than L2: 67 MB in + 67 MB out + 49 MB Fourier BSK + 61 MB KSK per step "
                             "(the BSK is meant to be L2-resident inside a launch)",
                "parallelism": f"independent blocks sharded over {world} GPU(s), keys replicated, no collective",
            },
            "verified_decrypt": all_ok,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel on this workload (4096 PBS),
                # from the ncu --set full capture in profiles/r1_final2_blind_rotate_ncu_full.csv (95.1 MB + 70.3 MB);
                # algorithmic bytes: 24.3 MB keyswitched inputs + 67.1 MB outputs + 48.6 MB Fourier BSK = 140.0 MB
                "traffic": 165.4e6 if B == BATCH else None, "traffic_unit": "bytes per launch",
                "kernel": "blind_rotate_kernel", "ms_per_launch": br_avg_ms,
                "flops_per_launch": (br_pbs / max(1, br_launches)) * FLOPS_PER_PBS,
                "peak_source": "DFMA microbenchmark measured in this run (fhestr_measure_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure; nominal 37.2 TFLOP/s",
                "keyswitch_ms_per_launch": ks_ms / max(1, br_launches),
                "kernel_share_of_step": br_ms / ms_total,
                # why the bound is FP64 and not HBM: the same kernel against the measured copy bandwidth
                "hbm": _hbm_side(165.4e6 if B == BATCH else None, br_avg_ms),
            },
            "e2e": {"value": e2e_value, "unit": "PBS/s", "h2d_bytes_per_step": int(B * 2049 * 8),
                    "d2h_bytes_per_step": int(B * 2049 * 8), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
        }
        if contains is not None:
            line["contains_256"] = contains
        if not args.no_cpu_baseline:
            v, cores, sample = cpu_port_rate(12.0)
            line["cpu_baseline"] = {"value": v, "unit": "PBS/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
