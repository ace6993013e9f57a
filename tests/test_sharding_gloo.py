"""Multi-rank host logic on CPU (gloo, world_size 2): every rank records the same graph, compiles it with
slot_align = world, computes ONLY its shard of each level's PBS jobs (fhestr_shard_range, the split
fhestr_program_run uses), all-gathers the level's result blocks in place, runs the replicated leveled jobs --
and must end with the same values as a single rank.  Plaintext block values stand in for ciphertexts."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from fhestring_b200.engine import load_library
        from fhestring_b200.graph import Graph
        from oracle import fhestring_plain as P
        from plain_exec import blocks_of, chars_of, run_jobs, shard_range

        lib = load_library()
        s = P.encrypt_str("the quick brown fox jumps over the lazy dog", 2)
        pat = [ord(c) for c in "lazy"]
        results = {}
        for method, ref in (("find", P.find(s, pat)), ("contains", P.contains(s, pat)), ("replace", None)):
            g = Graph()
            ids_s, slots_s = g.input_chars(len(s))
            ids_p, slots_p = g.input_chars(len(pat))
            args = [ids_s, ids_p] + ([ids_p] if method == "replace" else [])
            rs, rc = g.string_op(method, args, fast=True)
            outs = list(rs) if rs is not None else [rc]
            g.mark_output(outs)
            info = g.compile(world)
            jobs, off, npbs, first = g.program()
            luts = g.luts()
            values = np.full(info.slots_used, 31, np.int64)   # garbage everywhere a job has not written
            values[slots_s.reshape(-1)] = blocks_of(s).reshape(-1)
            values[slots_p.reshape(-1)] = blocks_of(pat).reshape(-1)
            ts, tv = g.trivials()
            values[ts.astype(np.int64)] = tv
            for l in range(info.n_levels):
                a, n = int(off[l]), int(npbs[l])
                lo, hi, per = shard_range(n, rank, world)
                clo, chi, cper = C.c_uint32(), C.c_uint32(), C.c_uint32()
                lib.fhestr_shard_range(C.c_uint32(n), C.c_uint32(rank), C.c_uint32(world), C.byref(clo), C.byref(chi), C.byref(cper))
                assert (lo, hi, per) == (clo.value, chi.value, cper.value)
                if n:
                    f = int(first[l])
                    # the engine requires: PBS job i of the level writes slot first + i, padded to per * world
                    assert [int(j["dst"]) for j in jobs[a:a + n]] == list(range(f, f + n))
                    assert f + per * world <= info.slots_used
                    run_jobs(values, jobs[a + lo:a + hi], luts)
                    mine = torch.from_numpy(values[f + rank * per:f + (rank + 1) * per].copy())
                    gathered = [torch.zeros_like(mine) for _ in range(world)]
                    dist.all_gather(gathered, mine)
                    values[f:f + per * world] = torch.cat(gathered).numpy()
                run_jobs(values, jobs[a + n:int(off[l + 1])], luts)   # leveled jobs: replicated
            got = chars_of(values, g.char_slots(outs))
            if ref is not None:
                assert int(got[0]) == ref, (method, got, ref)
            else:
                assert [int(v) for v in got] == P.replace(s, pat, pat)
            results[method] = [int(v) for v in got]
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_levels_and_agree(build_lib):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res[0] == res[1]


def test_shard_range_covers_every_job_once():
    sys.path.insert(0, HERE)
    from plain_exec import shard_range
    for n in (0, 1, 2, 7, 8, 9, 250, 8000):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_range(n, r, world)
                assert hi - lo <= per and per * world >= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
