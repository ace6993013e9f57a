"""Build libfhestr_engine.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfhestr_engine.so")
SOURCES = ["engine.cu", "blind_rotate.cu", "blind_rotate_wide.cu", "keyswitch.cu", "keyswitch_tc.cu", "aux_kernels.cu", "client.cpp", "graph.cpp", "strings.cpp", "graph_capi.cpp"]
HEADERS = ["kernels.cuh", "br_core.cuh", "br_wide.cuh", "fft32_gen.cuh", "graph.h", "strings.h", os.path.join("..", "..", "include", "fhestr_engine.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-Wno-stringop-overflow",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: tuple[str, ...] = ()) -> str:
    """variant="tag" with defines=("NAME=VALUE", ...) builds an experimental kernel variant (the FHESTR_BR_PREFETCH / FHESTR_BR_CVT_FP64 knobs
    of br_core.cuh) as libfhestr_engine_<tag>.so next to the default library; select it with FHESTR_ENGINE_LIB for
    A/B runs (scripts/ab_variants_gpu.sh)."""
    objdir = os.path.join(HERE, f"build_{variant}" if variant else "build")
    lib = os.path.join(HERE, f"libfhestr_engine_{variant}.so") if variant else LIB
    flags = NVCC_FLAGS + [f"-D{d}" for d in defines]
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src: str) -> str:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        if force or _stale(obj, [path] + hdrs):
            cmd = ["nvcc", *flags, "-x", "cu", "-c", path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(lib, objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs,
                               "-Xcompiler", "-fPIC", "-ldl"])
    return lib


if __name__ == "__main__":
    # usage: build.py [--force] [-v] [--variant TAG NAME=VALUE ...]
    tag, defs = "", ()
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        tag, defs = sys.argv[i + 1], tuple(a for a in sys.argv[i + 2:] if "=" in a)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=tag, defines=defs))
