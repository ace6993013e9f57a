"""Host-side mirror of the reference's public interface, over the engine's C ABI.

Same names, argument meaning and error behaviour as the reference so that parity tests read like its
own tests (/root/reference/src/main.rs:127-157):

    MyClientKey      /root/reference/src/client_key.rs:9-106
    MyServerKey      /root/reference/src/server_key/mod.rs:13-1876, trim.rs
    FheAsciiChar     /root/reference/src/ciphertext/fheasciichar.rs:7-168
    FheString        /root/reference/src/ciphertext/fhestring.rs:5-90
    FheStrip         /root/reference/src/ciphertext/fhestrip.rs:4-24

What differs is WHEN work happens: every FheAsciiChar method and every MyServerKey method RECORDS into the
op graph (fhestr_graph_*); decrypting (or MyServerKey.flush) compiles everything recorded so far into
dependency levels of independent PBS jobs and runs them on the GPU.  Nothing here computes on ciphertexts
in Python and there is no CPU path: without libfhestr_engine.so and a B200 the server key cannot be built.
"""
from __future__ import annotations

import numpy as np

from .client import ClientKey
from .engine import Engine, PARAM_MESSAGE_2_CARRY_2_KS_PBS
from .graph import Graph

MAX_BLOCKS = 4          # /root/reference/src/main.rs:23
STRING_PADDING = 1      # /root/reference/src/main.rs:12
MAX_REPETITIONS = 16    # /root/reference/src/main.rs:17
MAX_FIND_LENGTH = 255   # /root/reference/src/main.rs:20


class PublicParameters:
    """public_parameters.rs:4-17 -- carried everywhere by the reference and never used (fheasciichar.rs:22)"""

    def __init__(self, num_blocks: int = MAX_BLOCKS):
        self.num_blocks = num_blocks


class FheAsciiChar:
    """One encrypted u8 = 4 radix blocks.  Either a fresh host ciphertext (`ct`, from the client) or a value
    of a server key's graph (`sk`, `id`), or both once a fresh ciphertext has been handed to a server key."""

    __slots__ = ("ct", "sk", "id", "graph", "slots")

    def __init__(self, ct=None, sk=None, id=None, graph=None, slots=None):
        # `graph` is the Graph instance `id` belongs to: MyServerKey.reset() starts a new graph, and an id of the
        # old one must never be used in the new one (it would name an unrelated node).  `slots`: the four arena blocks
        # of a result computed by a cached plan (MyServerKey._run_plan): a value of `graph`'s lifetime without a node
        self.ct, self.sk, self.id, self.graph, self.slots = ct, sk, id, graph, slots

    @staticmethod
    def encrypt_trivial(value: int, public_parameters, server_key: "MyServerKey") -> "FheAsciiChar":  # :17-25
        return server_key._trivial(value)

    def _bin(self, op, server_key, other):
        return server_key._char_op(op, self, other)

    def eq(self, server_key, other): return self._bin("eq", server_key, other)          # :35
    def ne(self, server_key, other): return self._bin("ne", server_key, other)          # :40
    def le(self, server_key, other): return self._bin("le", server_key, other)          # :45
    def lt(self, server_key, other): return self._bin("lt", server_key, other)          # :50
    def ge(self, server_key, other): return self._bin("ge", server_key, other)          # :55
    def gt(self, server_key, other): return self._bin("gt", server_key, other)          # :60
    def bitand(self, server_key, other): return self._bin("bitand", server_key, other)  # :65
    def bitor(self, server_key, other): return self._bin("bitor", server_key, other)    # :74
    def sub(self, server_key, other): return self._bin("sub", server_key, other)        # :83
    def add(self, server_key, other): return self._bin("add", server_key, other)        # :87

    def if_then_else(self, server_key, true_value, false_value):                        # :93
        return server_key._char_op("if_then_else", self, true_value, false_value)

    def is_whitespace(self, server_key, public_parameters=None): return server_key._char_op("is_whitespace", self)  # :106
    def is_uppercase(self, server_key, public_parameters=None): return server_key._char_op("is_uppercase", self)    # :132
    def is_lowercase(self, server_key, public_parameters=None): return server_key._char_op("is_lowercase", self)    # :146
    def flip(self, server_key, public_parameters=None): return server_key._char_op("flip", self)                    # :161


class FheString:
    """fhestring.rs:5-9: Vec<FheAsciiChar> (+ the trivial constant 32 the reference keeps as `cst`)"""

    def __init__(self, chars):
        self.bytes = list(chars)

    @staticmethod
    def from_vec(chars, public_parameters=None, server_key=None):
        return FheString(chars)

    def __len__(self): return len(self.bytes)
    def __getitem__(self, i): return self.bytes[i]
    def __iter__(self): return iter(self.bytes)
    def push(self, c): self.bytes.append(c)


class FheStrip:
    """fhestrip.rs:4-7"""

    def __init__(self, string: FheString, pattern_found: FheAsciiChar):
        self.string, self.pattern_found = string, pattern_found

    @staticmethod
    def decrypt(strip: "FheStrip", client_key: "MyClientKey"):  # fhestrip.rs:15-23
        return client_key.decrypt(strip.string), client_key.decrypt_char(strip.pattern_found)


class FheSplit:
    """fhesplit.rs:5-8"""

    def __init__(self, buffers, pattern_found: FheAsciiChar):
        self.buffers, self.pattern_found = buffers, pattern_found

    @staticmethod
    def decrypt(fhe_split: "FheSplit", client_key: "MyClientKey"):  # fhesplit.rs:29-40
        chars = [c for b in fhe_split.buffers for c in b.bytes] + [fhe_split.pattern_found]
        vals = client_key.decrypt_padded(chars)     # one flush + one download for the whole result
        out, k = [], 0
        for b in fhe_split.buffers:
            raw = bytes(vals[k:k + len(b)])
            k += len(b)
            cut = raw.find(b"\0")
            out.append((raw if cut < 0 else raw[:cut]).decode("utf-8"))
        return out, vals[-1]


class MyClientKey:
    """client_key.rs:9-106.  Host-side keygen / encrypt / decrypt through fhestr_client_*."""

    def __init__(self, client: ClientKey, num_blocks: int = MAX_BLOCKS):
        assert num_blocks == MAX_BLOCKS
        self.client = client
        self.public_parameters = PublicParameters(num_blocks)
        self._server_keys = None

    @staticmethod
    def from_params(params: dict | None = None, num_blocks: int = MAX_BLOCKS, seed: int | None = None) -> "MyClientKey":  # :30-35
        """seed=None: OS entropy (like the reference); an explicit seed is for reproducible tests only"""
        prm = dict(PARAM_MESSAGE_2_CARRY_2_KS_PBS)
        prm.update(params or {})
        return MyClientKey(ClientKey(seed=seed, **prm), num_blocks)

    def get_server_key(self, **kw) -> "MyServerKey":  # :37-39
        if self._server_keys is None:
            self._server_keys = self.client.server_keys()
        p = self.client.params
        kw.setdefault("params", {f: getattr(p, f) for f, _ in p._fields_})
        return MyServerKey(self._server_keys[0], self._server_keys[1], **kw)

    def get_public_parameters(self) -> PublicParameters:  # :41-43
        return self.public_parameters

    @staticmethod
    def _check(string: str):
        assert all(ord(ch) < 128 and ch != "\0" for ch in string), \
            "The input string must only contain ascii letters and not include null characters"  # :52-55

    def encrypt(self, string: str, padding: int, public_parameters=None, server_key=None) -> FheString:  # :45-65
        self._check(string)
        cts = self.client.encrypt_u8((string + "\0" * padding).encode("ascii"))
        return FheString([FheAsciiChar(ct=cts[i]) for i in range(cts.shape[0])])

    def encrypt_no_padding(self, string: str):  # :67-79
        self._check(string)
        cts = self.client.encrypt_u8(string.encode("ascii"))
        return [FheAsciiChar(ct=cts[i]) for i in range(cts.shape[0])]

    def encrypt_char(self, plain_char: int) -> FheAsciiChar:  # :85-87
        return FheAsciiChar(ct=self.client.encrypt_u8(bytes([plain_char]))[0])

    def _host_cts(self, chars) -> np.ndarray:
        chars = list(chars)
        out = np.zeros((len(chars), 4, self.client.big), np.uint64)
        need = [i for i, c in enumerate(chars) if c.ct is None]
        for i, c in enumerate(chars):
            if c.ct is not None:
                out[i] = c.ct
        if need:
            sk = chars[need[0]].sk
            got = sk._download([chars[i] for i in need])
            for k, i in enumerate(need):
                out[i] = got[k]
        return out

    def decrypt_char(self, cipher_char: FheAsciiChar) -> int:  # :81-83
        return int(self.client.decrypt_u8(self._host_cts([cipher_char]))[0])

    def decrypt(self, cipher_string) -> str:  # :96-106 (truncates at the first NUL)
        b = bytes(self.client.decrypt_u8(self._host_cts(cipher_string)))
        cut = b.find(b"\0")
        return (b if cut < 0 else b[:cut]).decode("utf-8")

    def decrypt_padded(self, cipher_string) -> list[int]:
        """all chars, NULs included (what the plaintext oracle returns)"""
        return [int(v) for v in self.client.decrypt_u8(self._host_cts(cipher_string))]


class MyServerKey:
    """server_key/mod.rs:13-16 -- instead of holding a tfhe::integer::ServerKey it owns the engine (device
    key store + ciphertext arena) and the op graph.  `fast` selects the depth-minimised recording of the
    string methods (same plaintext for every input); fast=False issues the reference's own op order."""

    def __init__(self, bsk_std, ksk, params=None, device: int = 0, arena_blocks: int = 1 << 16, fast: bool = True,
                 rank: int = 0, world: int = 1, engine: Engine | None = None):
        self.engine = engine or Engine(arena_blocks=arena_blocks, device=device, **(params or {}))
        if engine is None:
            self.engine.load_keys(bsk_std, ksk)
        self.fast, self.rank, self.world = fast, rank, world
        self.key = self  # the reference passes `&my_server_key.key` around (main.rs:146)
        self.graph = Graph()
        self.last_info = None
        # plan cache: a string method called on fresh host ciphertexts right after reset() is recorded, compiled and
        # bound ONCE per (method, argument lengths, clear n, recording, world); the same call on new inputs re-runs
        # the bound program (upload, run) without recording anything.  plan_cache = False: always record.
        self.plan_cache = True
        self.plan_cache_size = 32       # bound programs kept (least recently used goes first; a 138 k-PBS program is 21 MB of device memory)
        self._plans = {}
        self._fresh = True
        self.plan_hits = 0

    def reset(self):
        """drop everything recorded and reuse the arena from slot 0 (one query = one graph)"""
        self.graph.close()
        self.graph = Graph()
        self._fresh = True

    # ---- plumbing between host ciphertexts, graph ids and the arena
    def _mine(self, c: FheAsciiChar) -> bool:
        return c.id is not None and c.sk is self and c.graph is self.graph

    def _plan_value(self, c: FheAsciiChar) -> bool:
        return c.slots is not None and c.sk is self and c.graph is self.graph

    def _materialise(self, chars):
        """results of a cached plan that are used as operands: fetch their ciphertexts, they re-enter as inputs"""
        need = [c for c in chars if c.ct is None and self._plan_value(c)]
        if need:
            got = self.engine.download_slots(np.concatenate([c.slots for c in need]).astype(np.int64)).reshape(len(need), 4, self.engine.big)
            for c, g in zip(need, got):
                c.ct = g

    def _adopt(self, c: FheAsciiChar) -> int:
        if self._mine(c):
            return c.id
        self._fresh = False
        self._materialise([c])
        if c.ct is None:
            raise ValueError("this FheAsciiChar is a value of another server key, or of this one before reset(), "
                             "and has no host ciphertext to re-upload")
        ids, slots = self.graph.input_chars(1)
        for b in range(4):
            self.engine.upload(int(slots[0, b]), c.ct[b])
        c.sk, c.id, c.graph = self, int(ids[0]), self.graph
        return c.id

    def _contiguous_run(self, chars):
        """[4 len(chars)][N+1] view over the chars' ciphertexts when they lie back to back in memory, else None"""
        big = self.engine.big
        first = chars[0].ct
        if not (isinstance(first, np.ndarray) and first.dtype == np.uint64 and first.flags.c_contiguous and first.shape == (4, big)):
            return None
        step = 4 * big * 8
        p0 = first.__array_interface__["data"][0]
        for i, c in enumerate(chars):
            a = c.ct
            if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags.c_contiguous and a.shape == (4, big)
                    and a.__array_interface__["data"][0] == p0 + i * step):
                return None
        owner = first
        while isinstance(owner.base, np.ndarray):
            owner = owner.base
        lo = owner.__array_interface__["data"][0]
        if owner.dtype != np.uint64 or not owner.flags.c_contiguous or p0 < lo or p0 + len(chars) * step > lo + owner.nbytes:
            return None
        k = (p0 - lo) // 8
        return owner.reshape(-1)[k:k + len(chars) * 4 * big].reshape(-1, big)

    def _adopt_all(self, chars):
        chars = list(chars)
        fresh = [c for c in chars if not self._mine(c)]
        self._fresh = False
        self._materialise(fresh)
        if any(c.ct is None for c in fresh):
            raise ValueError("an FheAsciiChar is a value of another server key, or of this one before reset(), "
                             "and has no host ciphertext to re-upload")
        if fresh:
            ids, slots = self.graph.input_chars(len(fresh))
            flat = slots.reshape(-1)
            # input slots are handed out consecutively: one upload
            assert (np.diff(flat.astype(np.int64)) == 1).all()
            run = self._contiguous_run(fresh)
            if run is not None:
                # the chars are consecutive views of one buffer (what MyClientKey.encrypt and any caller that keeps a
                # string's ciphertexts together hand over): upload it as it lies, no gathering copy on the host
                self.engine.upload(int(flat[0]), run)
            else:
                self.engine.upload(int(flat[0]), np.stack([c.ct for c in fresh]).reshape(-1, self.engine.big))
            for c, i in zip(fresh, ids):
                c.sk, c.id, c.graph = self, int(i), self.graph
        return np.array([c.id for c in chars], np.uint32)

    def _wrap(self, cid: int) -> FheAsciiChar:
        return FheAsciiChar(sk=self, id=int(cid), graph=self.graph)

    def _trivial(self, value: int) -> FheAsciiChar:
        self._fresh = False
        return self._wrap(self.graph.trivial_chars([value & 255])[0])

    def _char_op(self, op, a, b=None, c=None) -> FheAsciiChar:
        ids = [int(self._adopt(x)) for x in (a, b, c) if x is not None]
        return self._wrap(self.graph.char_op(op, *ids))

    def flush(self, outputs):
        """compile and run everything `outputs` depend on"""
        outputs = [c for c in outputs if not self._plan_value(c)]     # results of a cached plan exist already
        if not outputs:
            return
        if not all(self._mine(c) for c in outputs):
            raise ValueError("flush() of a char that is not a value of this server key's current graph")
        ids = np.array([c.id for c in outputs], np.uint32)
        self.graph.mark_output(ids)
        self.last_info = self.graph.compile(self.world)
        if self.last_info.slots_used > self.engine.arena_blocks:
            raise MemoryError(f"graph needs {self.last_info.slots_used} arena blocks, engine has {self.engine.arena_blocks}")
        self.graph.execute(self.engine, self.rank, self.world)

    def _download(self, chars) -> np.ndarray:
        chars = list(chars)
        self.flush(chars)
        slots = np.zeros((len(chars), 4), np.uint32)
        rec = [i for i, c in enumerate(chars) if not self._plan_value(c)]
        if rec:
            slots[rec] = self.graph.char_slots(np.array([chars[i].id for i in rec], np.uint32))
        for i, c in enumerate(chars):
            if self._plan_value(c):
                slots[i] = c.slots
        flat = slots.reshape(-1).astype(np.int64)
        if len(flat) and (np.diff(flat) == 1).all():
            out = self.engine.download(int(flat[0]), len(flat))
        else:
            out = self.engine.download_slots(flat)      # gathered on the device, one copy
        return out.reshape(len(chars), 4, self.engine.big)

    def _ids(self, a):
        """char ids of one argument: an FheString, a list of FheAsciiChar (unpadded pattern) or a single char"""
        chars = a.bytes if isinstance(a, FheString) else ([a] if isinstance(a, FheAsciiChar) else list(a))
        return self._adopt_all(chars) if chars else np.zeros(0, np.uint32)

    def _arg_chars(self, a):
        return a.bytes if isinstance(a, FheString) else ([a] if isinstance(a, FheAsciiChar) else list(a))

    def _str(self, method, *args, clear_n=0):
        key = None
        if self.plan_cache and self._fresh:
            lists = [self._arg_chars(a) for a in args]
            if all(len(l) for l in lists) and all(c.ct is not None and not self._mine(c) and not self._plan_value(c) for l in lists for c in l):
                key = (method, tuple(len(l) for l in lists), bool(self.fast), int(clear_n), int(self.world))
                plan = self._plans.pop(key, None)
                if plan is not None:
                    self._plans[key] = plan          # dicts keep insertion order: the most recently used plan is last
                    return self._run_plan(plan, lists)
        ids = [self._ids(a) for a in args]
        rs, rc = self.graph.string_op(method, ids, fast=self.fast, clear_n=clear_n)
        s = None if rs is None else FheString([self._wrap(i) for i in rs])
        c = None if rc is None else self._wrap(rc)
        if key is not None:
            self._make_plan(key, ids, s, c)
        return s, c

    def _make_plan(self, key, ids, s, c):
        """first call of a query shape: compute its results now (compile, bind, run, commit) and keep the bound program"""
        outs = (list(s.bytes) if s is not None else []) + ([c] if c is not None else [])
        oid = np.array([x.id for x in outs], np.uint32)
        self.graph.mark_output(oid)
        self.last_info = info = self.graph.compile(self.world)
        if info.slots_used > self.engine.arena_blocks:
            raise MemoryError(f"graph needs {info.slots_used} arena blocks, engine has {self.engine.arena_blocks}")
        prog = self.graph.bind(self.engine)
        triv = self.graph.trivials()
        prog.run(rank=self.rank, world=self.world)
        self.engine.sync()           # a failed run (peer barrier time-out) surfaces here, before anything is committed or cached
        self.graph.commit()
        oslots = self.graph.char_slots(oid)
        while len(self._plans) >= max(1, self.plan_cache_size):
            self._plans.pop(next(iter(self._plans)))["prog"].close()
        self._plans[key] = dict(
            prog=prog, in_first=[int(self.graph.char_slots(i[:1])[0, 0]) for i in ids], n_slots=int(info.slots_used), info=info,
            triv=tuple(a[np.argsort(triv[0], kind="stable")].copy() for a in triv), n_str=0 if s is None else len(s.bytes), has_char=c is not None, out_slots=oslots.copy())

    def _run_plan(self, plan, lists):
        """the same query shape on new inputs: upload them where the recording put its inputs, re-run the bound program"""
        for first, chars in zip(plan["in_first"], lists):
            run = self._contiguous_run(chars)
            self.engine.upload(first, run if run is not None else np.stack([c.ct for c in chars]).reshape(-1, self.engine.big))
        ts, tv = plan["triv"]                         # constants the program reads: another query may have used their blocks
        i = 0
        while i < len(ts):                            # one call per run of consecutive slots
            j = i + 1
            while j < len(ts) and int(ts[j]) == int(ts[j - 1]) + 1:
                j += 1
            self.engine.trivial(int(ts[i]), [int(v) for v in tv[i:j]])
            i = j
        plan["prog"].run(rank=self.rank, world=self.world)
        self.graph.reserve_slots(plan["n_slots"])
        self._fresh = False
        self.last_info = plan["info"]
        self.plan_hits += 1
        res = [FheAsciiChar(sk=self, graph=self.graph, slots=row.copy()) for row in plan["out_slots"]]
        s = FheString(res[:plan["n_str"]]) if plan["n_str"] else None
        c = res[plan["n_str"]] if plan["has_char"] else None
        return s, c

    def _clear(self, pattern: str):
        return [self._trivial(ord(ch)) for ch in pattern]

    # ---- the string methods (public_parameters is accepted and ignored, as in the reference)
    def to_upper(self, string, public_parameters=None): return self._str("to_upper", string)[0]            # mod.rs:65
    def to_lower(self, string, public_parameters=None): return self._str("to_lower", string)[0]            # mod.rs:110
    def contains(self, string, needle, public_parameters=None): return self._str("contains", string, needle)[1]        # :151
    def contains_clear(self, string, clear_needle, public_parameters=None): return self.contains(string, self._clear(clear_needle))  # :198
    def ends_with(self, string, needle, public_parameters=None): return self._str("ends_with", string, needle)[1]      # :241
    def ends_with_clear(self, string, clear_needle, public_parameters=None): return self.ends_with(string, self._clear(clear_needle))  # :303
    def starts_with(self, string, pattern, public_parameters=None): return self._str("starts_with", string, pattern)[1]  # :344
    def starts_with_clear(self, string, clear_pattern, public_parameters=None): return self.starts_with(string, self._clear(clear_pattern))  # :392
    def is_empty(self, string, public_parameters=None): return self._str("is_empty", string)[1]            # :431
    def len(self, string, public_parameters=None): return self._str("len", string)[1]                      # :478
    def repeat_clear(self, string, repetitions: int, public_parameters=None): return self._str("repeat_clear", string, clear_n=repetitions)[0]  # :517
    def repeat(self, string, repetitions, public_parameters=None): return self._str("repeat", string, repetitions)[0]  # :567
    def replace(self, string, from_, to, public_parameters=None): return self._str("replace", string, from_, to)[0]     # :624
    def replace_clear(self, string, clear_from, clear_to, public_parameters=None):                                      # :679
        return self.replace(string, self._clear(clear_from), self._clear(clear_to))
    def rfind(self, string, pattern, public_parameters=None): return self._find("rfind", string, pattern)  # :727
    def rfind_clear(self, string, clear_pattern, public_parameters=None): return self.rfind(string, self._clear(clear_pattern))  # :813
    def find(self, string, pattern, public_parameters=None): return self._find("find", string, pattern)    # :1010
    def find_clear(self, string, clear_pattern, public_parameters=None): return self.find(string, self._clear(clear_pattern))  # :1075
    def eq(self, string, other, public_parameters=None): return self._str("eq", string, other)[1]          # :1122
    def ne(self, string, other, public_parameters=None): return self._str("ne", string, other)[1]          # :1178
    def eq_ignore_case(self, string, other, public_parameters=None): return self._str("eq_ignore_case", string, other)[1]  # :1221
    def strip_prefix(self, string, pattern, public_parameters=None): return FheStrip(*self._str("strip_prefix", string, pattern))  # :1261
    def strip_suffix(self, string, pattern, public_parameters=None): return FheStrip(*self._str("strip_suffix", string, pattern))  # :1335
    def strip_prefix_clear(self, string, clear_pattern, public_parameters=None): return self.strip_prefix(string, self._clear(clear_pattern))  # :1421
    def strip_suffix_clear(self, string, clear_pattern, public_parameters=None): return self.strip_suffix(string, self._clear(clear_pattern))  # :1457
    def lt(self, string, other, public_parameters=None): return self._str("lt", string, other)[1]          # :1577
    def le(self, string, other, public_parameters=None): return self._str("le", string, other)[1]          # :1613
    def gt(self, string, other, public_parameters=None): return self._str("gt", string, other)[1]          # :1649
    def ge(self, string, other, public_parameters=None): return self._str("ge", string, other)[1]          # :1685
    def replacen(self, string, from_, to, n, public_parameters=None): return self._str("replacen", string, from_, to, n)[0]  # :1729
    def replacen_clear(self, string, clear_from, clear_to, clear_n: int, public_parameters=None):          # :1789
        return self.replacen(string, self._clear(clear_from), self._clear(clear_to), self._trivial(clear_n))
    def concatenate(self, string, other, public_parameters=None): return self._str("concatenate", string, other)[0]  # :1864
    def trim_end(self, string, public_parameters=None): return self._str("trim_end", string)[0]            # trim.rs:36
    def trim_start(self, string, public_parameters=None): return self._str("trim_start", string)[0]        # trim.rs:86
    def trim(self, string, public_parameters=None): return self._str("trim", string)[0]                    # trim.rs:146

    # ---- split family (server_key/split.rs)
    def _split(self, method, *args):
        ids = [self._ids(a) for a in args]
        bufs, found = self.graph.split_op(method, ids, fast=self.fast)
        return FheSplit([FheString([self._wrap(i) for i in row]) for row in bufs], self._wrap(found))

    def split(self, string, pattern, public_parameters=None): return self._split("split", string, pattern)                        # :1038
    def split_clear(self, string, clear_pattern, public_parameters=None): return self.split(string, self._clear(clear_pattern))      # :1094
    def rsplit(self, string, pattern, public_parameters=None): return self._split("rsplit", string, pattern)                      # :439
    def rsplit_clear(self, string, clear_pattern, public_parameters=None): return self.rsplit(string, self._clear(clear_pattern))    # :492
    def split_inclusive(self, string, pattern, public_parameters=None): return self._split("split_inclusive", string, pattern)    # :1155
    def split_inclusive_clear(self, string, clear_pattern, public_parameters=None): return self.split_inclusive(string, self._clear(clear_pattern))  # :1211
    def split_terminator(self, string, pattern, public_parameters=None): return self._split("split_terminator", string, pattern)  # :1267
    def split_terminator_clear(self, string, clear_pattern, public_parameters=None): return self.split_terminator(string, self._clear(clear_pattern))  # :1319
    def rsplit_terminator(self, string, pattern, public_parameters=None): return self._split("rsplit_terminator", string, pattern)  # :806
    def rsplit_terminator_clear(self, string, clear_pattern, public_parameters=None): return self.rsplit_terminator(string, self._clear(clear_pattern))  # :863
    def rsplit_once(self, string, pattern, public_parameters=None): return self._split("rsplit_once", string, pattern)            # :681
    def rsplit_once_clear(self, string, clear_pattern, public_parameters=None): return self.rsplit_once(string, self._clear(clear_pattern))  # :736
    def splitn(self, string, pattern, n, public_parameters=None): return self._split("splitn", string, pattern, n)                # :1497
    def splitn_clear(self, string, clear_pattern, clear_n: int, public_parameters=None): return self.splitn(string, self._clear(clear_pattern), self._trivial(clear_n))  # :1553
    def rsplitn(self, string, pattern, n, public_parameters=None): return self._split("rsplitn", string, pattern, n)              # :553
    def rsplitn_clear(self, string, clear_pattern, clear_n: int, public_parameters=None): return self.rsplitn(string, self._clear(clear_pattern), self._trivial(clear_n))  # :614
    def split_ascii_whitespace(self, string, public_parameters=None): return self._split("split_ascii_whitespace", string)        # :1377

    def _find(self, method, string, pattern):
        from .engine import EngineError
        try:
            return self._str(method, string, pattern)[1]
        except EngineError as ex:
            if "Maximum supported size for find reached" in str(ex):
                raise RuntimeError("Maximum supported size for find reached") from None  # the reference panics (mod.rs:743,1026)
            raise


def bubble_zeroes_right(server_key: MyServerKey, string: FheString) -> FheString:
    """utils.rs:28-46"""
    return server_key._str("bubble_zeroes_right", string)[0]
