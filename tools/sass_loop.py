"""Prints the instructions of the widest loop of a kernel (cuobjdump -sass), optionally filtered.
  python tools/sass_loop.py <obj> <kernel-regex> [filter-regex]"""
import re
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from sass_mix import load, parse

obj, pat = sys.argv[1], sys.argv[2]
flt = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
for name, body in load(obj, pat):
    rows = parse(body)
    best = None
    for addr, ins in rows:
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < addr:
            span = addr - int(m.group(1), 16)
            if best is None or span > best[0]:
                best = (span, int(m.group(1), 16), addr)
    print("==", name, hex(best[1]), hex(best[2]))
    for addr, ins in rows:
        if best[1] <= addr <= best[2] and (flt is None or flt.search(ins)):
            print("%05x  %s" % (addr, ins))
