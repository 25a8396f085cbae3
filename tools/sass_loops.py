#!/usr/bin/env python
"""Instruction budget of a kernel's loops, from the SASS in the built library.

    python tools/sass_loops.py [--lib aod_meh_hua_b200/libmehhua.so] [--kernel k2_dirichlet_kernelILb0] [--dump]

Runs `cuobjdump -sass`, splits the listing per kernel, finds every backward branch (a loop body =
[target, branch]) and prints, for the loops that contain MUFU instructions, the instruction count and
an opcode histogram grouped by pipe (XU = MUFU; FMA = FFMA/FMUL/FADD/IMAD; ALU = LOP3/SHF/IADD3/SEL/
FSEL/ISETP/FSETP/PRMT/...; LSU = LDS/STS/LDG/STG; other).  --dump also prints the SASS of the largest such loop
(the listing committed under profiles/).
"""
import argparse
import collections
import re
import subprocess
import sys

FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA32I", "FMUL32I", "FADD32I"}
XU = {"MUFU"}
LSU = {"LDS", "STS", "LDG", "STG", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "LDSM"}
ALU = {"LOP3", "SHF", "IADD3", "IADD", "SEL", "FSEL", "ISETP", "FSETP", "PRMT", "FMNMX", "PLOP3", "MOV", "LEA",
       "F2FP", "I2FP", "F2F", "I2F", "F2I", "VOTE", "POPC", "FLO", "IABS", "IMNMX", "BMSK", "SGXT", "LOP", "SHL", "SHR",
       "VIADD", "VIMNMX", "FCHK", "UMOV"}


def pipe(op: str) -> str:
    base = op.split(".")[0]
    if base in XU:
        return "XU"
    if base in FMA:
        return "FMA"
    if base in LSU:
        return "LSU"
    if base in ALU or base.startswith("U"):
        return "ALU"
    return "other"


def parse(lib: str):
    txt = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            addr = int(m.group(1), 16)
            ins = m.group(2).strip()
            pred = ""
            pm = re.match(r"(@!?U?P\d+)\s+(.*)", ins)
            if pm:
                pred, ins = pm.group(1), pm.group(2)
            kernels[cur].append((addr, pred, ins))
    return kernels


def loops(instrs):
    out = []
    for i, (addr, pred, ins) in enumerate(instrs):
        m = re.match(r"BRA(?:\.\S+)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                j = next(k for k, (a, _, _) in enumerate(instrs) if a == tgt)
                out.append((j, i))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="aod_meh_hua_b200/libmehhua.so")
    ap.add_argument("--kernel", default="k2_dirichlet_kernelILb0")
    ap.add_argument("--dump", action="store_true")
    args = ap.parse_args()
    ks = parse(args.lib)
    names = [k for k in ks if args.kernel in k]
    if not names:
        sys.exit(f"no kernel matching {args.kernel!r}; have: {sorted(ks)[:40]}")
    for name in names:
        instrs = ks[name]
        print(f"== {name}: {len(instrs)} SASS instructions")
        best = None
        for (j, i) in loops(instrs):
            body = instrs[j:i + 1]
            ops = [ins.split()[0] for (_, _, ins) in body]
            nmufu = sum(1 for o in ops if o.startswith("MUFU"))
            if nmufu == 0:
                continue
            hist = collections.Counter(pipe(o) for o in ops)
            byop = collections.Counter(o.split(".")[0] for o in ops)
            inner = any(j < jj and ii < i for (jj, ii) in loops(instrs))
            print(f"  loop 0x{instrs[j][0]:04x}..0x{instrs[i][0]:04x}: {len(body)} instr, MUFU {nmufu}, "
                  f"pipes {dict(hist)}{' (contains inner loops)' if inner else ''}")
            print("     ", ", ".join(f"{k} {v}" for k, v in byop.most_common(14)))
            if not inner and (best is None or len(body) > best[2]):
                best = (j, i, len(body))
        if args.dump and best:
            j, i, _ = best
            print(f"\n-- SASS of the largest MUFU-bearing innermost loop of {name}")
            for (addr, pred, ins) in instrs[j:i + 1]:
                print(f"  /*{addr:04x}*/ {pred:>6} {ins}")


if __name__ == "__main__":
    main()
