"""Decode the scheduling control bits of `cuobjdump -sass` output (Volta+ 128-bit encoding): stall count, yield, write / read
scoreboard index, wait mask.  usage: cuobjdump -sass -fun NAME lib.so | python tools/sass_ctrl.py [regex]"""
import re
import sys

pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
lines = sys.stdin.read().split("\n")
i = 0
n = 0
while i < len(lines):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = hi >> 41
            stall = ctrl & 0xF
            yld = (ctrl >> 4) & 1
            wbar = (ctrl >> 5) & 7
            rbar = (ctrl >> 8) & 7
            wait = (ctrl >> 11) & 0x3F
            txt = m.group(2).strip()
            if pat is None or pat.search(txt):
                print("%5d %s  st%-2d %s W%s R%s wait[%s]  %s" % (n, m.group(1), stall, "Y" if yld else "-", wbar if wbar != 7 else "-", rbar if rbar != 7 else "-",
                                                           "".join(str(b) for b in range(6) if wait >> b & 1), txt[:90]))
            n += 1
            i += 2
            continue
    i += 1
