#!/usr/bin/env python3
"""Instruction statistics of the loops of a kernel from `cuobjdump -sass`: every backward branch closes a loop;
prints, per loop (outermost first), its size in instructions / bytes and an opcode-class histogram.
usage: sass_loop_stats.py file.sass [kernel-name-substring]"""
import collections
import re
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    kern, ins = None, {}
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            ins[kern] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and kern:
            ins[kern].append((int(m.group(1), 16), m.group(2).strip()))
    for k, lst in ins.items():
        if want not in k:
            continue
        print(f"== {k}: {len(lst)} instructions, {16 * len(lst)} bytes")
        loops = []
        for addr, txt in lst:
            m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", txt)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
        loops.sort(key=lambda ab: ab[0] - ab[1])
        for lo, hi in loops[:6]:
            body = [t for a, t in lst if lo <= a <= hi]
            h = collections.Counter()
            for t in body:
                op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
                base = op.split(".")[0]
                cls = ("FP64" if base in ("DFMA", "DADD", "DMUL") else
                       "LDS/STS" if base in ("LDS", "STS") else
                       "LDG" if base in ("LDG", "LD") else
                       "local" if base in ("LDL", "STL") else
                       "CVT" if base in ("I2F", "F2I", "F2F") else
                       "MOV" if base in ("MOV", "UMOV", "IMAD") and (".MOV" in op or base != "IMAD") else
                       "sync" if base in ("BAR", "WARPSYNC", "BSSY", "BSYNC", "NOP", "BRA", "SYNCS") else "int/other")
                h[cls] += 1
            print(f"  loop 0x{lo:x}..0x{hi:x}: {len(body)} instr = {16 * len(body) / 1024:.1f} KB  " +
                  "  ".join(f"{c}:{n}" for c, n in h.most_common()))


main()
