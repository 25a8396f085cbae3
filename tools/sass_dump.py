#!/usr/bin/env python
"""Print the SASS of one address range of one kernel:  python tools/sass_dump.py KERNEL_SUBSTR 0xSTART 0xEND [lib]"""
import re
import subprocess
import sys

kern, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
lib = sys.argv[4] if len(sys.argv) > 4 else "aod_meh_hua_b200/libmehhua.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
on = False
for line in txt.splitlines():
    if "Function :" in line:
        on = kern in line
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and lo <= int(m.group(1), 16) <= hi:
        print(m.group(1), m.group(2))
