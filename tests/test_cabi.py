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
