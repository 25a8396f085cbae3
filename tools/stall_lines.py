#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one kernel of an .ncu-rep, attributed to SOURCE LINES.
`ncu --page source --csv` gives per-SASS-address counters; `nvdisasm -g` of the library's cubin gives the line of every
address; this joins the two (no GPU needed).

    python tools/stall_lines.py REPORT.ncu-rep KERNEL_SUBSTR [TOP_N] [MANGLED_SUBSTR]

KERNEL_SUBSTR selects the kernel in the report (demangled name); MANGLED_SUBSTR (default: the same) selects the function in
the disassembly - give the mangled instantiation for templates, e.g. k1c_gather_kernelILi81ELi1E."""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
mangled = sys.argv[4] if len(sys.argv) > 4 else kern
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(root, "aod_meh_hua_b200", "libmehhua.so")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
sec, hdr, samp, base = None, None, {}, None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        sec = kern in r[1] and not samp
        continue
    if r and r[0] == "Address":
        hdr = r
        ia, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        base = None
        continue
    if sec and hdr and len(r) == len(hdr):
        a = int(r[0], 16)
        base = a if base is None else base
        samp[a - base] = (int(r[ia]), int(r[ie]))

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
line, infn, amap, taken = None, False, {}, False
for ln in dis.splitlines():
    if ln.lstrip().startswith(".text."):
        infn = (mangled in ln) and not taken
        taken = taken or infn
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        amap[int(m.group(1), 16)] = line
agg = {}
for a, (s, e) in samp.items():
    d = agg.setdefault(amap.get(a), [0, 0])
    d[0] += s
    d[1] += e
tot, tote = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"{kern}: {tot} stall samples, {tote} warp instructions")
for k, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        try:
            text = open(k[0]).read().splitlines()[k[1] - 1].strip()[:100]
        except OSError:
            pass
    where = f"{os.path.basename(k[0])}:{k[1]}" if k else "?"
    print(f"{100 * s / tot:5.1f}% of stalls, {100 * e / tote:5.1f}% of instructions  {where}  {text}")
