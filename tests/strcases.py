"""Shared description of the string-method test cases: how each method's arguments are encrypted by the
reference's tests / CLI (padded string, unpadded pattern, encrypted u8) and how its result is decoded."""
from __future__ import annotations

import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

# method -> (argument kinds, result kind).  "s" padded string, "p" unpadded pattern (encrypt_no_padding),
# "n" encrypted u8, "c" clear integer.  Result: "u8", "str", "strip" (string, found flag)
SIGNATURES = {
    "contains": ("sp", "u8"), "ends_with": ("sp", "u8"), "starts_with": ("sp", "u8"),
    "find": ("sp", "u8"), "rfind": ("sp", "u8"),
    "is_empty": ("s", "u8"), "len": ("s", "u8"),
    "to_upper": ("s", "str"), "to_lower": ("s", "str"),
    "trim": ("s", "str"), "trim_start": ("s", "str"), "trim_end": ("s", "str"),
    "repeat": ("sn", "str"), "repeat_clear": ("sc", "str"),
    "replace": ("spp", "str"), "replacen": ("sppn", "str"),
    "eq": ("ss", "u8"), "ne": ("ss", "u8"), "eq_ignore_case": ("ss", "u8"),
    "lt": ("ss", "u8"), "le": ("ss", "u8"), "gt": ("ss", "u8"), "ge": ("ss", "u8"),
    "concatenate": ("ss", "str"),
    "strip_prefix": ("sp", "strip"), "strip_suffix": ("sp", "strip"),
    # split family: result kind "split" = (list of padded buffers, found flag)
    "split": ("sp", "split"), "rsplit": ("sp", "split"), "split_inclusive": ("sp", "split"),
    "split_terminator": ("sp", "split"), "rsplit_terminator": ("sp", "split"), "rsplit_once": ("sp", "split"),
    "splitn": ("spn", "split"), "rsplitn": ("spn", "split"), "split_ascii_whitespace": ("s", "split"),
}


def reference_cases():
    with open(os.path.join(HERE, "golden", "reference_tests.json")) as f:
        return json.load(f)


def encode_args(method, args, padding):
    """plaintext arguments -> lists of u8 / ints, the way the reference's tests encrypt them"""
    kinds, _ = SIGNATURES[method]
    out = []
    for kind, a in zip(kinds, args):
        if kind == "s":
            out.append([ord(ch) for ch in a] + [0] * padding)
        elif kind == "p":
            out.append([ord(ch) for ch in a])
        else:
            out.append(int(a))
    return out


def cut_at_nul(chars):
    out = []
    for v in chars:
        if v == 0:
            break
        out.append(chr(v))
    return "".join(out)


def decode_result(method, res):
    kind = SIGNATURES[method][1]
    if kind == "u8":
        return int(res)
    if kind == "str":
        return cut_at_nul(res)
    if kind == "split":   # FheSplit::decrypt + trim_vector (utils.rs:59-70); the reference's tests ignore the flag
        bufs = [cut_at_nul(b) for b in res[0]]
        while bufs and bufs[0] == "":
            bufs.pop(0)
        while bufs and bufs[-1] == "":
            bufs.pop()
        return bufs
    return [cut_at_nul(res[0]), int(res[1])]
