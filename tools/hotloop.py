#!/usr/bin/env python3
"""Development aid: print the SASS of the innermost loops (back-edges) of a kernel in a .so with a scalar-equivalent issue count.
usage: tools/hotloop.py lib.so kernel_substring"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in blk.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:P\d,\s*)?(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", t)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                loops.append((addr[tgt], i))
    print("==", name, len(ins), "instructions;", len(loops), "back-edges")
    for (b, e) in loops:
        body = ins[b:e + 1]
        if len(body) > 120 or len(body) < 12:
            continue
        packed = sum(1 for _, t in body if re.search(r"\b(FFMA2|FADD2|FMUL2)\b", t))
        mufu = sum(1 for _, t in body if "MUFU" in t)
        print("-- loop %04x..%04x: %d instr, %d packed, %d MUFU -> scalar-equivalent issue %d" % (body[0][0], body[-1][0], len(body), packed, mufu, len(body) + packed))
        if "-v" in sys.argv:
            for a, t in body:
                print("   %04x %s" % (a, t))
