"""ctypes binding of libfhestr_engine.so (include/fhestr_engine.h).

This is plumbing only: every call goes through the C ABI that the reference's Rust `-sys` crate would
bind.  There is no Python or CPU implementation of any PBS step behind it -- if the
shared library is missing or no B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FHESTR_ENGINE_LIB") or os.path.join(_HERE, "libfhestr_engine.so")  # override: A/B runs of two builds
MAX_TERMS = 16


class EngineError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("N", C.c_int32), ("k", C.c_int32),
        ("pbs_base_log", C.c_int32), ("pbs_level", C.c_int32),
        ("ks_base_log", C.c_int32), ("ks_level", C.c_int32), ("delta_log", C.c_int32),
    ]


class Job(C.Structure):
    _fields_ = [
        ("dst", C.c_uint32), ("lut", C.c_int32), ("n_terms", C.c_uint32),
        ("src", C.c_uint32 * MAX_TERMS), ("coeff", C.c_int32 * MAX_TERMS),
        ("constant", C.c_uint64),
    ]


JOB_DTYPE = np.dtype(
    [("dst", "<u4"), ("lut", "<i4"), ("n_terms", "<u4"), ("src", "<u4", (MAX_TERMS,)),
     ("coeff", "<i4", (MAX_TERMS,)), ("_pad", "<u4"), ("constant", "<u8")]
)
assert JOB_DTYPE.itemsize == C.sizeof(Job), (JOB_DTYPE.itemsize, C.sizeof(Job))

# PARAM_MESSAGE_2_CARRY_2_KS_PBS (reference: src/main.rs:3,43; values: SURVEY.md A.1)
PARAM_MESSAGE_2_CARRY_2_KS_PBS = dict(
    n=742, N=2048, k=1, pbs_base_log=23, pbs_level=1, ks_base_log=3, ks_level=5, delta_log=59
)


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -m fhestring_b200.build` "
            "(there is no fallback implementation)"
        )
    lib = C.CDLL(LIB_PATH)
    lib.fhestr_last_error.restype = C.c_char_p
    lib.fhestr_last_error.argtypes = [C.c_void_p]
    lib.fhestr_arena_ptr.restype = C.c_void_p
    lib.fhestr_arena_ptr.argtypes = [C.c_void_p]
    lib.fhestr_kernel_launches.restype = C.c_uint64
    lib.fhestr_kernel_launches.argtypes = [C.c_void_p]
    lib.fhestr_engine_destroy.restype = None
    lib.fhestr_engine_destroy.argtypes = [C.c_void_p]
    lib.fhestr_program_destroy.restype = None
    lib.fhestr_program_destroy.argtypes = [C.c_void_p]
    lib.fhestr_shard_range.restype = None
    return lib


def make_jobs(n: int) -> np.ndarray:
    return np.zeros(n, JOB_DTYPE)


def _u64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


class Program:
    def __init__(self, engine: "Engine", handle, n_levels: int, level_sizes):
        self.engine, self.handle, self.n_levels, self.level_sizes = engine, handle, n_levels, list(level_sizes)

    def run(self, first_level=0, last_level=None, rank=0, world=1):
        last_level = self.n_levels if last_level is None else last_level
        self.engine._ck(self.engine.lib.fhestr_program_run(
            self.engine.h, self.handle, C.c_uint32(first_level), C.c_uint32(last_level),
            C.c_uint32(rank), C.c_uint32(world)))

    def close(self):
        if self.handle:
            self.engine.lib.fhestr_program_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Device-resident key store + ciphertext arena + batched PBS, one per process and GPU."""

    def __init__(self, arena_blocks: int, device: int = 0, external_arena_ptr: int | None = None, **params):
        self.lib = load_library()
        prm = dict(PARAM_MESSAGE_2_CARRY_2_KS_PBS)
        prm.update(params)
        self.params = Params(**prm)
        self.N, self.n = self.params.N, self.params.n
        self.big = self.N + 1
        self.arena_blocks = arena_blocks
        h = C.c_void_p()
        rc = self.lib.fhestr_engine_create(C.byref(self.params), C.c_int(device), C.c_uint64(arena_blocks),
                                           C.c_void_p(external_arena_ptr or 0), C.byref(h))
        if rc != 0:
            raise EngineError(f"fhestr_engine_create failed ({rc}): {self.lib.fhestr_last_error(None).decode()}")
        self.h = h
        self._luts: dict[tuple, int] = {}

    # -- plumbing
    def _ck(self, rc: int):
        if rc != 0:
            raise EngineError(f"engine call failed ({rc}): {self.lib.fhestr_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhestr_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self.lib.fhestr_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def sync(self):
        self._ck(self.lib.fhestr_sync(self.h))

    def arena_ptr(self) -> int:
        return int(self.lib.fhestr_arena_ptr(self.h))

    def kernel_launches(self) -> int:
        return int(self.lib.fhestr_kernel_launches(self.h))

    def set_br_mode(self, mode: int, wide_max_jobs: int = 0):
        """0 = kernel by level size (default), 1 = throughput kernel only, 2 = latency kernel only"""
        self._ck(self.lib.fhestr_set_br_mode(self.h, C.c_int(mode), C.c_int(wide_max_jobs)))

    def set_keyswitch_path(self, path: int):
        """0 = tensor cores (tcgen05 kind::i8 limb-split GEMM, default), 1 = CUDA cores"""
        self._ck(self.lib.fhestr_set_keyswitch_path(self.h, C.c_int(path)))

    def set_timing(self, enable: bool):
        self._ck(self.lib.fhestr_set_timing(self.h, C.c_int(1 if enable else 0)))

    def get_timing(self):
        """(keyswitch_ms, blind_rotate_ms, blind_rotate_launches, blind_rotate_pbs) since the last call"""
        ks, br = C.c_double(), C.c_double()
        nl, npbs = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.fhestr_get_timing(self.h, C.byref(ks), C.byref(br), C.byref(nl), C.byref(npbs)))
        return ks.value, br.value, int(nl.value), int(npbs.value)

    # -- multi-GPU: one process per GPU, the engine's own NCCL communicator (fhestr_comm_*)
    def comm_init(self, rank: int, world: int):
        """rank 0 creates the NCCL unique id, torch.distributed (any backend) carries it to the other ranks"""
        import torch
        import torch.distributed as dist
        uid = np.zeros(128, np.uint8)
        if rank == 0:
            self._ck(self.lib.fhestr_comm_unique_id(uid.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(uid)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        uid = np.ascontiguousarray(t.cpu().numpy())
        self._ck(self.lib.fhestr_comm_init(self.h, C.c_uint32(rank), C.c_uint32(world), uid.ctypes.data_as(C.c_void_p)))
        self.rank, self.world = rank, world

    def peer_attach(self, rank: int, world: int):
        """map every rank's arena with cudaIpc (handles carried by torch.distributed): results are stored into all
        arenas by the blind-rotation epilogue, no collective on the data path"""
        import torch
        import torch.distributed as dist
        mine = np.zeros(128, np.uint8)
        self._ck(self.lib.fhestr_peer_export(self.h, mine[:64].ctypes.data_as(C.c_void_p), mine[64:].ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(mine)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        allh = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        allh = np.stack([x.cpu().numpy() for x in allh])
        ah, fh = np.ascontiguousarray(allh[:, :64]), np.ascontiguousarray(allh[:, 64:])
        self._ck(self.lib.fhestr_peer_attach(self.h, C.c_uint32(rank), C.c_uint32(world), ah.ctypes.data_as(C.c_void_p),
                                             fh.ctypes.data_as(C.c_void_p)))
        self.rank, self.world = rank, world
        dist.barrier()

    def peer_detach(self):
        self._ck(self.lib.fhestr_peer_detach(self.h))

    def peer_timed_out(self) -> bool:
        v = C.c_uint32()
        self._ck(self.lib.fhestr_peer_status(self.h, C.byref(v)))
        return bool(v.value)

    def comm_destroy(self):
        self._ck(self.lib.fhestr_comm_destroy(self.h))

    def shard_range(self, n_jobs: int, rank: int, world: int):
        lo, hi, per = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.lib.fhestr_shard_range(C.c_uint32(n_jobs), C.c_uint32(rank), C.c_uint32(world), C.byref(lo), C.byref(hi), C.byref(per))
        return lo.value, hi.value, per.value

    # -- keys / LUTs
    def load_keys(self, bsk_std: np.ndarray, ksk: np.ndarray):
        bsk_std = np.ascontiguousarray(bsk_std, np.uint64)
        ksk = np.ascontiguousarray(ksk, np.uint64)
        assert bsk_std.size == self.n * 4 * self.N, "bsk must be [n][1][2][2][N]"
        assert ksk.size == self.N * self.params.ks_level * (self.n + 1), "ksk must be [N][ks_level][n+1]"
        self._ck(self.lib.fhestr_load_keys(self.h, _u64p(bsk_std), _u64p(ksk)))

    def lut(self, table) -> int:
        key = tuple(int(x) for x in table)
        if key not in self._luts:
            t = np.asarray(key, np.uint8)
            assert len(t) == 1 << (63 - self.params.delta_log)
            out = C.c_int32()
            self._ck(self.lib.fhestr_lut_register(self.h, t.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(out)))
            self._luts[key] = out.value
        return self._luts[key]

    def lut_download(self, lut_id: int) -> np.ndarray:
        out = np.zeros(self.N, np.uint64)
        self._ck(self.lib.fhestr_lut_download(self.h, C.c_int32(lut_id), _u64p(out)))
        return out

    # -- arena
    def upload(self, first_block: int, cts: np.ndarray):
        cts = np.ascontiguousarray(cts, np.uint64).reshape(-1, self.big)
        self._ck(self.lib.fhestr_ct_upload(self.h, C.c_uint32(first_block), C.c_uint32(cts.shape[0]), _u64p(cts)))
        self._keep = cts  # pageable source must outlive the async copy
        self.sync()

    def download(self, first_block: int, count: int) -> np.ndarray:
        out = np.zeros((count, self.big), np.uint64)
        self._ck(self.lib.fhestr_ct_download(self.h, C.c_uint32(first_block), C.c_uint32(count), _u64p(out)))
        return out

    def download_slots(self, slots) -> np.ndarray:
        """arena blocks named one by one -> [len(slots)][N+1]: gathered on the device, one copy"""
        s = np.ascontiguousarray(slots, np.uint32).ravel()
        out = np.zeros((len(s), self.big), np.uint64)
        if len(s):
            self._ck(self.lib.fhestr_ct_download_slots(self.h, s.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint32(len(s)), _u64p(out)))
        return out

    def trivial(self, first_block: int, values):
        v = np.ascontiguousarray(values, np.uint8)
        self._ck(self.lib.fhestr_ct_trivial(self.h, C.c_uint32(first_block), C.c_uint32(len(v)),
                                            v.ctypes.data_as(C.POINTER(C.c_uint8))))

    # -- hot path
    def pbs_batch(self, jobs: np.ndarray):
        jobs = np.ascontiguousarray(jobs, JOB_DTYPE)
        self._ck(self.lib.fhestr_pbs_batch(self.h, jobs.ctypes.data_as(C.POINTER(Job)), C.c_uint32(len(jobs))))

    def program(self, jobs: np.ndarray, level_offsets) -> Program:
        jobs = np.ascontiguousarray(jobs, JOB_DTYPE)
        off = np.ascontiguousarray(level_offsets, np.uint32)
        h = C.c_void_p()
        self._ck(self.lib.fhestr_program_create(self.h, jobs.ctypes.data_as(C.POINTER(Job)),
                                                off.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                C.c_uint32(len(off) - 1), C.byref(h)))
        return Program(self, h, len(off) - 1, np.diff(off))

    # -- test / measurement hooks
    def debug_keyswitch(self, jobs: np.ndarray) -> np.ndarray:
        jobs = np.ascontiguousarray(jobs, JOB_DTYPE)
        out = np.zeros((len(jobs), self.n + 1), np.uint64)
        self._ck(self.lib.fhestr_debug_keyswitch(self.h, jobs.ctypes.data_as(C.POINTER(Job)),
                                                 C.c_uint32(len(jobs)), _u64p(out)))
        return out

    def debug_blind_rotate(self, ks: np.ndarray, lut_ids=None, init_acc: np.ndarray | None = None) -> np.ndarray:
        ks = np.ascontiguousarray(ks, np.uint64).reshape(-1, self.n + 1)
        count = ks.shape[0]
        out = np.zeros((count, 2, self.N), np.uint64)
        ids = None if lut_ids is None else np.ascontiguousarray(lut_ids, np.int32)
        init = None if init_acc is None else np.ascontiguousarray(init_acc, np.uint64)
        self._ck(self.lib.fhestr_debug_blind_rotate(
            self.h, _u64p(ks), None if ids is None else ids.ctypes.data_as(C.POINTER(C.c_int32)),
            None if init is None else _u64p(init), C.c_uint32(count), _u64p(out)))
        return out

    def measure_fp64_peak(self) -> tuple[float, float]:
        tf, mhz = C.c_double(), C.c_double()
        self._ck(self.lib.fhestr_measure_fp64_peak(self.h, C.byref(tf), C.byref(mhz)))
        return tf.value, mhz.value


def single_term_jobs(dst, src, lut) -> np.ndarray:
    """jobs arena[dst[i]] = PBS_lut[i](arena[src[i]])"""
    dst = np.asarray(dst)
    jobs = make_jobs(len(dst))
    jobs["dst"] = dst
    jobs["lut"] = lut
    jobs["n_terms"] = 1
    jobs["src"][:, 0] = src
    jobs["coeff"][:, 0] = 1
    return jobs

