#!/usr/bin/env python
"""Print one function's SASS (address + instruction) from a cuobjdump -sass dump.
usage: sass_func.py dump.sass <substring of mangled name> [start_hex end_hex]"""
import re, sys
txt = open(sys.argv[1]).read()
for f in re.split(r'\n\s+Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    if sys.argv[2] in name:
        ins = re.findall(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', f)
        lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
        hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
        print(name, len(ins), 'instructions')
        for a, i in ins:
            if lo <= int(a, 16) <= hi:
                print(a, i[:100])
        break
