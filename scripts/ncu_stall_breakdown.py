"""Per-phase warp-stall breakdown of the blind-rotation CMUX loop from an `ncu --set full --import-source on` report.
usage: ncu -i REPORT.ncu-rep --page source --csv --print-source=sass > sass.csv ; python scripts/ncu_stall_breakdown.py sass.csv
Prints: stall reasons over the loop, samples per instruction family, and a position profile (the loop cut into bins
of equal instruction count) -- the tables in profiles/r1_br_stall_breakdown.md."""
import sys

import numpy as np
import pandas as pd


def family(op: str) -> str:
    op = str(op)
    if op.startswith(("DFMA", "DADD", "DMUL")):
        return "FP64"
    if op.startswith(("LDS", "STS")):
        return "SMEM"
    if op.startswith("LDG"):
        return "LDG"
    if op.startswith(("I2F", "F2I", "F2F")):
        return "CVT"
    if op.startswith(("BAR", "WARPSYNC", "BSYNC", "BSSY", "NANOSLEEP")):
        return "SYNC"
    return "INT"


def main(path: str, bins: int = 48) -> None:
    df = pd.read_csv(path, skiprows=1)
    df["op"] = (df["Source"].str.strip().str.replace(r"^@!?U?P\d+\s+", "", regex=True).str.split().str[0].str.rstrip(";"))
    df["samples"] = df["# Samples"]
    hot = df["Instructions Executed"].max()
    loop = df[df["Instructions Executed"] > 0.9 * hot].reset_index(drop=True).copy()
    stall_cols = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
    total = loop["samples"].sum()
    print(f"loop instructions {len(loop)}, executed {int(loop['Instructions Executed'].median())} times each, "
          f"{total} samples ({100 * total / df['samples'].sum():.1f} % of the kernel)")
    s = loop[stall_cols].sum().sort_values(ascending=False)
    print("\nstall reason, % of loop samples")
    for k, v in s.head(10).items():
        print(f"  {k.replace('stall_', ''):14s} {100 * v / total:5.1f}")
    loop["fam"] = loop["op"].map(family)
    g = loop.groupby("fam").agg(instructions=("op", "size"), samples=("samples", "sum"))
    g["pct"] = 100 * g["samples"] / total
    g["samples_per_instr"] = g["samples"] / g["instructions"]
    print("\nby instruction family\n", g.round(1).to_string())
    for f in ("FP64", "SMEM", "INT", "LDG", "CVT"):
        sub = loop[loop["fam"] == f][stall_cols].sum().sort_values(ascending=False)
        print(f"  {f}: " + ", ".join(f"{k.replace('stall_', '')} {100 * v / max(1, loop[loop['fam'] == f]['samples'].sum()):.0f}%"
                                   for k, v in sub.head(6).items()))
    loop["bin"] = np.arange(len(loop)) * bins // len(loop)
    pv = loop.pivot_table(index="bin", columns="fam", values="samples", aggfunc="sum", fill_value=0)
    pv["total"] = pv.sum(axis=1)
    pv["pct"] = (100 * pv["total"] / total).round(1)
    cnt = loop.pivot_table(index="bin", columns="fam", values="op", aggfunc="size", fill_value=0)
    print("\nposition profile (samples per bin | instruction counts per bin)\n", pv.join(cnt, rsuffix="_n").to_string())


if __name__ == "__main__":
    main(sys.argv[1])
