#!/usr/bin/env python
"""SASS evidence for profiles/ (runs here, no GPU): opcode histogram of one kernel of an object file + its hottest loop.

  python tools/sass_summary.py fss_b200/csrc/build/inst_k1_p0_s0.o 'point_kernelILi0ELi0ELi0ELi5E' > profiles/r02_sass_point.txt

The listing is `cuobjdump -sass` of the object the shipped library is linked from.  The "main loop" is the largest
backward-branch region of the kernel (the level loop of the point / gen kernels, the depth-first walk of EvalAll)."""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    body = next((f for f in funcs[1:] if pat in f.split("\n", 1)[0]), None)
    if body is None:
        sys.exit(f"no kernel matching {pat!r} in {obj}")
    name = body.split("\n", 1)[0].strip()
    ins = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))

    def opcode(s):
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        return s.split()[0]

    hist = collections.Counter(opcode(s) for _, s in ins)
    print(f"# {name}\n# object: {obj}\n# {len(ins)} instructions\n")
    print("## opcode histogram (whole kernel, static)\n")
    for op, n in hist.most_common():
        print(f"{n:6d}  {op}")
    fam = collections.Counter()
    for op, n in hist.items():
        fam[op.split(".")[0]] += n
    print("\n## by mnemonic family\n")
    for op, n in fam.most_common(24):
        print(f"{n:6d}  {op}")
    marks = ("UTMALDG", "UTMASTG", "SYNCS", "LDS", "STS", "LDG", "STG", "PRMT", "IDP", "LOP3", "LDCU", "UTMAPF", "UBLKCP")
    print("\n## instructions that carry the design\n")
    for k in marks:
        print(f"{sum(n for op, n in hist.items() if op.startswith(k)):6d}  {k}*")
    # largest backward branch = main loop
    best = None
    for i, (addr, s) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+.*?(0x[0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    if best:
        loop = [(a, s) for a, s in ins if best[0] <= a <= best[1]]
        lh = collections.Counter(opcode(s).split(".")[0] for _, s in loop)
        print(f"\n## main loop: {len(loop)} instructions ({best[0]:#x} .. {best[1]:#x}), by family\n")
        for op, n in lh.most_common(16):
            print(f"{n:6d}  {op}")
        print("\n## main loop listing (first 120 and last 20 instructions)\n")
        show = loop[:120] + ([(None, "...")] if len(loop) > 140 else []) + (loop[-20:] if len(loop) > 140 else loop[120:])
        for a, s in show:
            print(("        " if a is None else f"/*{a:05x}*/ ") + s)


if __name__ == "__main__":
    main()
