"""ctypes binding of the op-graph half of the C ABI (fhestr_graph_*, include/fhestr_engine.h).

The graph records the reference's per-char primitives (/root/reference/src/ciphertext/fheasciichar.rs)
and string methods (/root/reference/src/server_key/*.rs) as radix-block PBS jobs and compiles them into
dependency levels; the engine executes the levels.  Recording and compiling are host-only, so this
module works without a GPU (tests interpret the compiled job list on plaintext values); executing needs
the engine.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import EngineError, Job, JOB_DTYPE, load_library

CHAR_OPS = {
    "eq": 0, "ne": 1, "le": 2, "lt": 3, "ge": 4, "gt": 5, "bitand": 6, "bitor": 7, "sub": 8, "add": 9,
    "if_then_else": 10, "is_whitespace": 11, "is_uppercase": 12, "is_lowercase": 13, "flip": 14,
}
METHODS = {
    "contains": 0, "ends_with": 1, "starts_with": 2, "is_empty": 3, "len": 4, "repeat_clear": 5, "repeat": 6,
    "replace": 7, "rfind": 8, "find": 9, "eq": 10, "ne": 11, "eq_ignore_case": 12, "strip_prefix": 13,
    "strip_suffix": 14, "lt": 15, "le": 16, "gt": 17, "ge": 18, "replacen": 19, "concatenate": 20,
    "to_upper": 21, "to_lower": 22, "trim_end": 23, "trim_start": 24, "trim": 25, "bubble_zeroes_right": 26,
}


SPLIT_METHODS = {
    "split": 0, "rsplit": 1, "split_inclusive": 2, "split_terminator": 3, "rsplit_terminator": 4, "rsplit_once": 5,
    "splitn": 6, "rsplitn": 7, "split_ascii_whitespace": 8,
}


class StrArg(C.Structure):
    _fields_ = [("chars", C.POINTER(C.c_uint32)), ("len", C.c_uint32)]


class GraphInfo(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("n_jobs", C.c_uint32), ("n_luts", C.c_uint32),
                ("n_trivial", C.c_uint32), ("slots_used", C.c_uint32), ("n_pbs", C.c_uint64),
                ("n_pbs_recorded", C.c_uint64)]


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


class Graph:
    def __init__(self, delta_log: int = 59):
        self.lib = load_library()
        self.lib.fhestr_graph_last_error.restype = C.c_char_p
        self.lib.fhestr_graph_last_error.argtypes = [C.c_void_p]
        self.lib.fhestr_graph_destroy.restype = None
        self.lib.fhestr_graph_destroy.argtypes = [C.c_void_p]
        h = C.c_void_p()
        rc = self.lib.fhestr_graph_create(C.c_int32(delta_log), C.byref(h))
        if rc:
            raise EngineError(f"fhestr_graph_create failed ({rc})")
        self.h = h
        self.info: GraphInfo | None = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhestr_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise EngineError(f"graph call failed ({rc}): {self.lib.fhestr_graph_last_error(self.h).decode()}")

    # ---- recording
    def input_chars(self, count: int):
        """-> (char ids [count], arena slots [count][4]) of freshly encrypted chars (FheAsciiChar::encrypt)"""
        ids = np.zeros(count, np.uint32)
        slots = np.zeros((count, 4), np.uint32)
        self._ck(self.lib.fhestr_graph_input_chars(self.h, C.c_uint32(count), _u32p(ids), _u32p(slots)))
        return ids, slots

    def trivial_chars(self, values):
        v = np.ascontiguousarray(values, np.uint8)
        ids = np.zeros(len(v), np.uint32)
        self._ck(self.lib.fhestr_graph_trivial_chars(self.h, v.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                     C.c_uint32(len(v)), _u32p(ids)))
        return ids

    def char_op(self, op: str, a: int, b: int = 0, c: int = 0) -> int:
        out = C.c_uint32()
        self._ck(self.lib.fhestr_graph_char_op(self.h, C.c_int(CHAR_OPS[op]), C.c_uint32(int(a)), C.c_uint32(int(b)),
                                               C.c_uint32(int(c)), C.byref(out)))
        return out.value

    def string_op(self, method: str, args, fast: bool = True, clear_n: int = 0):
        """args: sequences of char ids (an encrypted u8 argument is a 1-char sequence).
        -> (string result as an array of char ids or None, char result id or None)"""
        arrs = [np.ascontiguousarray(a, np.uint32).ravel() for a in args]
        sa = (StrArg * max(1, len(arrs)))()
        for i, a in enumerate(arrs):
            sa[i].chars = _u32p(a)
            sa[i].len = len(a)
        cap = 4096
        for _ in range(2):
            out = np.zeros(cap, np.uint32)
            out_len, out_char = C.c_uint32(0), C.c_uint32(0xFFFFFFFF)
            rc = self.lib.fhestr_graph_string_op(self.h, C.c_int(METHODS[method]), C.c_int(1 if fast else 0), sa,
                                                 C.c_uint32(len(arrs)), C.c_uint64(int(clear_n)), _u32p(out),
                                                 C.c_uint32(cap), C.byref(out_len), C.byref(out_char))
            if rc and out_len.value > cap:
                cap = out_len.value
                continue
            self._ck(rc)
            break
        kind_str = method in ("repeat_clear", "repeat", "replace", "replacen", "concatenate", "to_upper", "to_lower",
                              "trim_end", "trim_start", "trim", "bubble_zeroes_right", "strip_prefix", "strip_suffix")
        s = out[:out_len.value].copy() if kind_str else None
        ch = None if out_char.value == 0xFFFFFFFF else out_char.value
        return s, ch

    def split_op(self, method: str, args, fast: bool = True):
        """the split family (split.rs) -> (buffers as an array [n_buffers][buffer_len] of char ids, found char id)"""
        arrs = [np.ascontiguousarray(a, np.uint32).ravel() for a in args]
        sa = (StrArg * max(1, len(arrs)))()
        for i, a in enumerate(arrs):
            sa[i].chars = _u32p(a)
            sa[i].len = len(a)
        L = len(arrs[0]) + 2
        cap = L * (L + 1)
        out = np.zeros(cap, np.uint32)
        nb, bl, found = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self.lib.fhestr_graph_split_op(self.h, C.c_int(SPLIT_METHODS[method]), C.c_int(1 if fast else 0), sa,
                                                C.c_uint32(len(arrs)), _u32p(out), C.c_uint32(cap), C.byref(nb),
                                                C.byref(bl), C.byref(found)))
        return out[:nb.value * bl.value].reshape(nb.value, bl.value).copy(), found.value

    def mark_output(self, ids):
        a = np.ascontiguousarray(ids, np.uint32).ravel()
        self._ck(self.lib.fhestr_graph_mark_output(self.h, _u32p(a), C.c_uint32(len(a))))

    # ---- compile / inspect
    def compile(self, slot_align: int = 1) -> GraphInfo:
        info = GraphInfo()
        self._ck(self.lib.fhestr_graph_compile(self.h, C.c_uint32(slot_align), C.byref(info)))
        self.info = info
        return info

    def program(self):
        """-> (jobs, level_offsets, level_pbs, level_first_dst) of the last compile"""
        i = self.info
        jobs = np.zeros(max(1, i.n_jobs), JOB_DTYPE)
        off = np.zeros(i.n_levels + 1, np.uint32)
        npbs = np.zeros(max(1, i.n_levels), np.uint32)
        first = np.zeros(max(1, i.n_levels), np.uint32)
        self._ck(self.lib.fhestr_graph_get_program(self.h, jobs.ctypes.data_as(C.POINTER(Job)), _u32p(off),
                                                   _u32p(npbs), _u32p(first)))
        return jobs[:i.n_jobs], off, npbs[:i.n_levels], first[:i.n_levels]

    def luts(self) -> np.ndarray:
        t = np.zeros((max(1, self.info.n_luts), 16), np.uint8)
        self._ck(self.lib.fhestr_graph_get_luts(self.h, t.ctypes.data_as(C.POINTER(C.c_uint8))))
        return t[:self.info.n_luts]

    def trivials(self):
        n = self.info.n_trivial
        slots, vals = np.zeros(max(1, n), np.uint32), np.zeros(max(1, n), np.uint8)
        self._ck(self.lib.fhestr_graph_get_trivials(self.h, _u32p(slots), vals.ctypes.data_as(C.POINTER(C.c_uint8))))
        return slots[:n], vals[:n]

    def char_slots(self, ids) -> np.ndarray:
        a = np.ascontiguousarray(ids, np.uint32).ravel()
        out = np.zeros((len(a), 4), np.uint32)
        self._ck(self.lib.fhestr_graph_char_slots(self.h, _u32p(a), C.c_uint32(len(a)), _u32p(out)))
        return out

    def reserve_slots(self, first_free: int):
        """arena blocks [0, first_free) are in use by a bound program run outside this graph (plan cache)"""
        self._ck(self.lib.fhestr_graph_reserve_slots(self.h, C.c_uint32(int(first_free))))

    # ---- run on the engine
    def execute(self, engine, rank: int = 0, world: int = 1):
        self._ck(self.lib.fhestr_graph_execute(self.h, engine.h, C.c_uint32(rank), C.c_uint32(world)))

    def bind(self, engine):
        """-> engine Program of the last compile (LUTs registered, trivial outputs written); pair with commit()"""
        from .engine import Program
        h = C.c_void_p()
        self._ck(self.lib.fhestr_graph_bind(self.h, engine.h, C.byref(h)))
        _, off, _, _ = self.program()
        return Program(engine, h, len(off) - 1, np.diff(off))

    def commit(self):
        self._ck(self.lib.fhestr_graph_commit(self.h))
