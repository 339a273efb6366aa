#!/usr/bin/env python
"""Decode the scheduling control bits of cuobjdump -sass output (sm_100a: stall, yield, write/read scoreboard, wait mask).

    cuobjdump -sass kernels.cu.o | python tools/sass_ctl.py [first_line last_line]
"""
import re
import sys

lines = sys.stdin.read().splitlines()
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
hi = int(sys.argv[2]) if len(sys.argv) > 2 else len(lines)
pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* 0x([0-9a-f]{16}) \*/")
pat2 = re.compile(r"^\s*/\* 0x([0-9a-f]{16}) \*/")
i = 0
while i < len(lines) - 1:
    m = pat.search(lines[i])
    m2 = pat2.search(lines[i + 1]) if m else None
    if m and m2:
        if lo <= i <= hi:
            w = int(m2.group(1), 16)
            ctl = w >> 41
            stall = ctl & 0xf; yld = (ctl >> 4) & 1; wbar = (ctl >> 5) & 7; rbar = (ctl >> 8) & 7; wait = (ctl >> 11) & 0x3f
            print("%6d %s  st%-2d %s w%s r%s wait[%s]  %s" % (i, m.group(1), stall, "Y" if not yld else " ", wbar if wbar < 7 else "-", rbar if rbar < 7 else "-",
                                                           "".join(str(b) for b in range(6) if wait >> b & 1), m.group(2).strip()))
        i += 2
    else:
        i += 1
