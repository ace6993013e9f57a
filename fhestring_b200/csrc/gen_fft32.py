#!/usr/bin/env python3
"""Generate fft32_gen.cuh: the straight-line, register-resident 32-point transform of the four-step
negacyclic FFT (1024 complex points = 32 x 32) used by the blind-rotation kernel, and the table of its
twist along n1.

The transform is a Cooley-Tukey network written as polynomial evaluation: the 32 inputs are the
coefficients of c(x) and the outputs are c(x_k) at the 32 roots of x^32 = rho.  One stage reduces
c modulo (x^m - r) and (x^m + r) with r = sqrt(rho), i.e. butterflies (a, b) -> (a + r b, a - r b)
with ONE twiddle per sub-block; a general butterfly is 6 FMAs (a - r b is formed as 2a - (a + r b)),
a trivial one (r = 1, i) is 4 add/sub.  All twiddles are literals, so they live in the constant
bank and no twiddle registers are needed.

  rho = 1: plain DFT-32 at exp(-2 pi i k / 32), 388 ops -- the ONE codelet of the kernel's rolled loop: all four
  passes of a CMUX step run it (the inverse passes on swapped real / imaginary parts), and the twist exp(i pi n1/64)
  of the folded transform is applied outside it from kFft32Twist*.  (Round 1 generated four networks, two of them with
  the twist merged in: 480 + 388 + 388 + 512 ops, 57.7 KB of straight-line loop.)

Outputs are returned in NATURAL index order (the permutation is free: pure register renaming).
The script checks the emitted network numerically against the dense DFT matrix before writing.

Run:  python gen_fft32.py   (writes fft32_gen.cuh next to this file; the output is committed)
"""
import cmath
import math
import os

import numpy as np

R = 32


class Prog:
    """tiny SSA builder; values are names, constants are python floats"""

    def __init__(self):
        self.lines = []
        self.cnt = 0
        self.ops = 0

    def new(self, expr):
        name = f"t{self.cnt}"
        self.cnt += 1
        self.lines.append((name, expr))
        self.ops += 1
        return name


TWIST = [f(math.pi * n / 64) for n in range(R) for f in (math.cos, math.sin)]
CONST_POOL = []   # distinct |twiddle| values; referenced as FFTC(i) so that DFMA reads them from the constant bank


def lit(x):
    """literal for a twiddle: pooled by absolute value (values closer than 4e-16 are the same twiddle), sign kept"""
    x = float(x)
    if x in (0.0, 1.0, -1.0, 2.0, -2.0):
        return repr(x)
    a = abs(x)
    for i, c in enumerate(CONST_POOL):
        if abs(c - a) < 4e-16:
            return ("-" if x < 0 else "") + f"FFTC({i})"
    CONST_POOL.append(a)
    return ("-" if x < 0 else "") + f"FFTC({len(CONST_POOL) - 1})"


def emit_network(prog, re, im, rho_angle):
    """re/im: lists of 32 SSA names (coefficients). Returns list of (angle, re_name, im_name) leaves."""

    def rec(cr, ci, angle):
        m = len(cr)
        if m == 1:
            return [(angle, cr[0], ci[0])]
        h = m // 2
        ra = angle / 2.0  # r = exp(i*ra), r^2 = rho
        wr, wi = math.cos(ra), math.sin(ra)
        # snap trivial twiddles
        triv = None
        for name, (tr, ti) in {"one": (1, 0), "i": (0, 1), "mone": (-1, 0), "mi": (0, -1)}.items():
            if abs(wr - tr) < 1e-15 and abs(wi - ti) < 1e-15:
                triv = name
        lo_r, lo_i, hi_r, hi_i = [], [], [], []
        for j in range(h):
            ar, ai, br, bi = cr[j], ci[j], cr[j + h], ci[j + h]
            if triv == "one":
                pr = prog.new(f"{ar} + {br}"); pi = prog.new(f"{ai} + {bi}")
                qr = prog.new(f"{ar} - {br}"); qi = prog.new(f"{ai} - {bi}")
            elif triv == "i":  # r*b = (-bi, br)
                pr = prog.new(f"{ar} - {bi}"); pi = prog.new(f"{ai} + {br}")
                qr = prog.new(f"{ar} + {bi}"); qi = prog.new(f"{ai} - {br}")
            elif triv == "mone":
                pr = prog.new(f"{ar} - {br}"); pi = prog.new(f"{ai} - {bi}")
                qr = prog.new(f"{ar} + {br}"); qi = prog.new(f"{ai} + {bi}")
            elif triv == "mi":  # r*b = (bi, -br)
                pr = prog.new(f"{ar} + {bi}"); pi = prog.new(f"{ai} - {br}")
                qr = prog.new(f"{ar} - {bi}"); qi = prog.new(f"{ai} + {br}")
            else:
                x = prog.new(f"fma({lit(wr)}, {br}, {ar})")
                pr = prog.new(f"fma({lit(-wi)}, {bi}, {x})")
                y = prog.new(f"fma({lit(wr)}, {bi}, {ai})")
                pi = prog.new(f"fma({lit(wi)}, {br}, {y})")
                qr = prog.new(f"fma(2.0, {ar}, -{pr})")
                qi = prog.new(f"fma(2.0, {ai}, -{pi})")
            lo_r.append(pr); lo_i.append(pi); hi_r.append(qr); hi_i.append(qi)
        return rec(lo_r, lo_i, ra) + rec(hi_r, hi_i, ra + math.pi)

    return rec(re, im, rho_angle)


def build(name, rho_angle, root_angle_of_k, post_angle_of_k=None, doc=""):
    """root_angle_of_k(k): angle of the evaluation point that must land in output slot k.
    post_angle_of_k(k): optional constant rotation applied to output k afterwards."""
    prog = Prog()
    re = [f"re[{j}]" for j in range(R)]
    im = [f"im[{j}]" for j in range(R)]
    leaves = emit_network(prog, re, im, rho_angle)
    # map leaves to natural outputs
    outs = [None] * R
    for ang, lr, li in leaves:
        found = None
        for k in range(R):
            d = (ang - root_angle_of_k(k)) / (2 * math.pi)
            if abs(d - round(d)) < 1e-9:
                found = k
        assert found is not None and outs[found] is None, (name, ang)
        outs[found] = (lr, li)
    final = []
    for k in range(R):
        lr, li = outs[k]
        if post_angle_of_k is not None:
            a = post_angle_of_k(k)
            c, s = math.cos(a), math.sin(a)
            if abs(c - 1) < 1e-15 and abs(s) < 1e-15:
                pass
            else:
                x = prog.new(f"{lit(c)} * {lr}")
                nr = prog.new(f"fma({lit(-s)}, {li}, {x})")
                y = prog.new(f"{lit(c)} * {li}")
                ni = prog.new(f"fma({lit(s)}, {lr}, {y})")
                lr, li = nr, ni
        final.append((lr, li))
    body = [f"// {doc}", f"// {prog.ops} FP64 ops per call",
            f"FHE_HD void {name}(double (&re)[32], double (&im)[32]) {{"]
    for n, e in prog.lines:
        body.append(f"    const double {n} = {e};")
    for k, (lr, li) in enumerate(final):
        body.append(f"    re[{k}] = {lr}; im[{k}] = {li};")
    body.append("}")
    return "\n".join(body), prog, final


def simulate(prog, final, x):
    """numerically run the SSA program on complex input vector x (len 32)"""
    env = {}
    for j in range(R):
        env[f"re[{j}]"] = x[j].real
        env[f"im[{j}]"] = x[j].imag

    import re as _re

    def ev(expr):
        # replace names by values
        def sub(m):
            return repr(float(env[m.group(0)]))
        e = _re.sub(r"re\[\d+\]|im\[\d+\]|t\d+", sub, expr)
        return eval(e, {"fma": lambda a, b, c: a * b + c, "FFTC": lambda i: CONST_POOL[i]})

    for n, e in prog.lines:
        env[n] = ev(e)
    return np.array([complex(env[a], env[b]) for a, b in final])


def main():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(R) + 1j * rng.standard_normal(R)
    funcs = []
    specs = [
        ("fft32_dft", 0.0, lambda k: -2 * math.pi * k / 32, None,
         "plain DFT-32: out[k] = sum_n in[n] * exp(-2*pi*i*k*n/32)"),
    ]
    for name, rho, root, post, doc in specs:
        src, prog, final = build(name, rho, root, post, doc)
        got = simulate(prog, final, x)
        want = np.array([sum(x[n] * cmath.exp(1j * root(k) * n) for n in range(R)) for k in range(R)])
        if post is not None:
            want = want * np.array([cmath.exp(1j * post(k)) for k in range(R)])
        err = np.max(np.abs(got - want))
        assert err < 1e-12, (name, err)
        print(f"{name}: {prog.ops} ops, max err {err:.2e}")
        funcs.append(src)
    hdr = [
        "// GENERATED by gen_fft32.py -- do not edit.  The 32-point register-resident transform of the",
        "// four-step (32 x 32) negacyclic FFT of the blind-rotation kernel and its twist table.  See gen_fft32.py.",
        "#pragma once",
        "#ifndef FHE_HD",
        "#ifdef __CUDACC__",
        "#define FHE_HD __host__ __device__ __forceinline__",
        "#else",
        "#define FHE_HD inline",
        "#endif",
        "#endif",
        "#include <math.h>",
        "namespace fhestr {",
        "// twiddle constants: in the constant bank on the device (DFMA takes c[bank][offset] operands directly; as",
        "// literals every use cost two UMOVs), a plain table on the host",
        "#ifdef __CUDACC__",
        "static __constant__ double kFft32ConstDev[] = {" + ", ".join(repr(c) for c in CONST_POOL) + "};",
        "#endif",
        "static const double kFft32ConstHost[] = {" + ", ".join(repr(c) for c in CONST_POOL) + "};",
        "// twist of the folded negacyclic transform along n1: cos, sin of pi n1 / 64 (br_core.cuh applies it outside the",
        "// shared plain DFT-32 instead of merging it into a second network)",
        "#ifdef __CUDACC__",
        "static __constant__ double kFft32TwistDev[] = {" + ", ".join(repr(v) for v in TWIST) + "};",
        "#endif",
        "static const double kFft32TwistHost[] = {" + ", ".join(repr(v) for v in TWIST) + "};",
        "#ifdef __CUDA_ARCH__",
        "#define FFTC(i) kFft32ConstDev[i]",
        "#else",
        "#define FFTC(i) kFft32ConstHost[i]",
        "#endif",
    ]
    out = "\n".join(hdr) + "\n\n" + "\n\n".join(funcs) + "\n\n#undef FFTC\n} // namespace fhestr\n"
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fft32_gen.cuh")
    with open(path, "w") as f:
        f.write(out)
    print("wrote", path, len(out), "bytes")


if __name__ == "__main__":
    main()
