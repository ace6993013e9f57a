"""The C-ABI library loads, exports every symbol include/fhestr_engine.h declares, its job struct has
the layout the Python/Rust bindings assume, and (on a box without a GPU) it refuses to create an
engine instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, _has_gpu


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fhestr_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fhestr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(build_lib):
    lib = C.CDLL(build_lib)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/fhestr_engine.h but not exported"


def test_job_struct_layout():
    from fhestring_b200.engine import JOB_DTYPE, Job, MAX_TERMS
    assert C.sizeof(Job) == JOB_DTYPE.itemsize == 4 + 4 + 4 + 4 * MAX_TERMS * 2 + 4 + 8
    assert Job.constant.offset == JOB_DTYPE.fields["constant"][1]
    assert Job.coeff.offset == JOB_DTYPE.fields["coeff"][1]


def test_rejects_unsupported_parameters(build_lib):
    from fhestring_b200.engine import Engine, EngineError
    with pytest.raises(EngineError, match="unsupported parameter set"):
        Engine(arena_blocks=4, N=1024)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu(build_lib):
    from fhestring_b200.engine import Engine, EngineError
    with pytest.raises(EngineError, match="no usable CUDA device|no CPU fallback"):
        Engine(arena_blocks=4)


# which Blackwell units each kernel is built on, read from the shipped library's SASS (cuobjdump is part of the toolkit
# that built it): a rebuild that silently drops the tensor-memory twiddles, the tcgen05 GEMM or the bulk-TMA key ring
# still passes every numerical test, so the instruction mix is pinned here
SASS_FEATURES = {
    "blind_rotate_kernel": ["LDTM", "STTM", "UTCATOMSWS", "DFMA"],            # tcgen05.ld / st / alloc, FP64 FMA
    "ks_gemm_tc_kernel": ["UTCIMMA", "UTMALDG", "UTCBAR", "LDTM", "SYNCS"],   # tcgen05.mma kind::i8, TMA, tcgen05.commit
    "blind_rotate_wide_kernel": ["UBLKCP", "SYNCS", "DFMA"],                  # bulk TMA key tiles + mbarriers
    "blind_rotate_wide2_kernel": ["UBLKCP", "SYNCS", "DFMA"],
}


def test_kernels_use_the_blackwell_units_they_claim(build_lib):
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", build_lib], capture_output=True, text=True, check=True).stdout
    bodies, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            bodies[name] = []
        elif name:
            bodies[name].append(line)
    for kernel, mnemonics in SASS_FEATURES.items():
        hits = [k for k in bodies if kernel + "E" in k or kernel + "I" in k]
        assert hits, f"{kernel} not found in the library's SASS"
        text = "\n".join(bodies[hits[0]])
        for mn in mnemonics:
            assert re.search(r"\b" + mn, text), f"{kernel}: no {mn} instruction in its SASS"
        assert "sm_100a" in sass or "sm_100" in sass
