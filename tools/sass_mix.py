"""Instruction mix of the main loop of a kernel, read from `cuobjdump -sass` (no GPU needed).

  python tools/sass_mix.py <object-or-so> <mangled-kernel-name-substring> [--all-loops]

The main loop is taken to be the widest backward branch.  Everything between its target and the
branch is counted once, i.e. both sides of the in-loop branches (sincos slow paths, resync) are
included; --exclude-calls drops nothing since CALL targets sit outside the loop.
"""
import collections
import re
import subprocess
import sys


def load(obj, pat):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    out = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if re.search(pat, name):
            out.append((name, f))
    return out


INS = re.compile(r"^\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);")


def parse(body):
    rows = []
    for ln in body.split("\n"):
        m = INS.match(ln)
        if m:
            addr = int(m.group(1), 16)
            ins = m.group(2).strip()
            rows.append((addr, ins))
    return rows


def family(op, operands=""):
    op0 = op.split(".")[0]
    if op0 == "DFMA":
        # B200 measurement (tools/microbench2.cu): a DFMA reading three vector registers issues every
        # 3 cycles per SM sub-partition, one with a uniform/constant/immediate operand every ~2.1
        srcs = [o.strip() for o in operands.split(",")][1:]
        nreg = sum(1 for o in srcs if re.match(r"^-?\|?R\d+", o))
        return "fp64:DFMA.%dreg" % nreg
    if op0 in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        return "fp64:" + op0
    if op0 in ("MUFU",):
        return "mufu:" + op
    if op0.startswith("LDS") or op0.startswith("STS"):
        return "smem:" + op0 + ("." + ".".join(op.split(".")[1:]) if "." in op else "")
    if op0.startswith("LDG") or op0.startswith("STG") or op0.startswith("LDL") or op0.startswith("STL") \
            or op0.startswith("LDC") or op0.startswith("ATOM") or op0.startswith("RED"):
        return "mem:" + op0
    if op0 in ("SHFL",):
        return "shfl"
    if op0 in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "NOP", "BREAK"):
        return "ctrl:" + op0
    return "other:" + op0


def reuse_adjusted(rows, lo, hi):
    """FP64 issue cycles of [lo, hi] with the measured B200 costs: DFMA with three vector-register
    sources 3.1 cycles unless one of them sits in the operand-reuse cache (same register, same slot,
    flagged .reuse by the preceding FP64 instruction), everything else ~2.1."""
    cyc, n3, n3r = 0.0, 0, 0
    prev = {}
    for addr, ins in rows:
        if not (lo <= addr <= hi):
            continue
        parts = ins.split()
        k = 1 if parts[0].startswith("@") else 0
        op0 = parts[k].split(".")[0]
        if op0 not in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
            continue
        srcs = [o.strip() for o in " ".join(parts[k + 1:]).split(",")][1:]
        if op0 == "DSETP":
            srcs = srcs[-3:]
        regs = {}
        for slot, o in enumerate(srcs):
            m = re.match(r"^-?\|?(R\d+)", o)
            if m:
                regs[slot] = (m.group(1), ".reuse" in o)
        if op0 == "DFMA" and len(regs) == 3:
            n3 += 1
            hit = any(prev.get(slot) == r for slot, (r, _) in regs.items())
            n3r += hit
            cyc += 2.2 if hit else 3.1
        else:
            cyc += 2.1
        prev = {slot: r for slot, (r, flag) in regs.items() if flag}
    return cyc, n3, n3r


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    for name, body in load(obj, pat):
        rows = parse(body)
        loops = []
        for addr, ins in rows:
            m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", ins)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr:
                    loops.append((addr - tgt, tgt, addr))
        loops.sort(reverse=True)
        print("==", name, "(%d instructions, %d backward branches)" % (len(rows), len(loops)))
        for span, lo, hi in (loops if "--all-loops" in sys.argv else loops[:1]):
            cnt = collections.Counter()
            for addr, ins in rows:
                if lo <= addr <= hi:
                    parts = ins.split()
                    k = 1 if parts[0].startswith("@") else 0
                    cnt[family(parts[k], " ".join(parts[k + 1:]))] += 1
            tot = sum(cnt.values())
            fp64 = sum(v for k, v in cnt.items() if k.startswith("fp64"))
            cyc = sum(v * (3.0 if k.endswith("3reg") else 2.1 if "DFMA" in k else 2.0)
                      for k, v in cnt.items() if k.startswith("fp64"))
            print("  loop 0x%x..0x%x: %d instructions, %d fp64 (%.0f%%), fp64 pipe cost %.0f cycles/iteration"
                  % (lo, hi, tot, fp64, 100.0 * fp64 / tot, cyc))
            c2, n3, n3r = reuse_adjusted(rows, lo, hi)
            print("  with operand reuse: %.0f cycles (%d of %d three-register DFMAs follow a .reuse of the same operand)"
                  % (c2, n3r, n3))
            for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:28]:
                print("    %-28s %5d" % (k, v))


if __name__ == "__main__":
    main()
