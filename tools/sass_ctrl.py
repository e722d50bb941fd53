#!/usr/bin/env python
"""Annotate one function's SASS with the Volta+ control fields decoded from the 128-bit encoding:
stall count, yield, write/read scoreboard slot, wait mask.
usage: cuobjdump -sass lib.so > dump.sass ; sass_ctrl.py dump.sass <mangled-name substring> [lo_hex hi_hex]"""
import re, sys
txt = open(sys.argv[1]).read()
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
for f in re.split(r'\n\s+Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[2] not in name:
        continue
    lines = f.split('\n')
    print(name)
    i = 0
    while i < len(lines):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/', lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
            if m2:
                addr = int(m.group(1), 16)
                hiw = int(m2.group(1), 16)
                stall = (hiw >> 41) & 0xf
                yld = (hiw >> 45) & 1
                wbar = (hiw >> 46) & 7
                rbar = (hiw >> 49) & 7
                wait = (hiw >> 52) & 0x3f
                if lo <= addr <= hi:
                    w = ''.join(str(k) for k in range(6) if wait >> k & 1) or '-'
                    print(f"{m.group(1)} st{stall:2d} {'Y' if yld else ' '} W{wbar if wbar != 7 else '-'} R{rbar if rbar != 7 else '-'} wait[{w:6s}] {m.group(2)[:90]}")
                i += 2
                continue
        i += 1
    break
