#!/usr/bin/env python
"""Static SASS accounting of the force kernel's loops: for every backward branch of the selected function,
the opcode histogram of the loop body [target, branch].  Usage: sass_loops.py [lib.so] [function substring]."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gplum_b200/libgplum_b200.so"
fun = sys.argv[2] if len(sys.argv) > 2 else "force_pass_kernelILi2"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if fun not in name:
        continue
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr_index = {a: k for k, (a, _) in enumerate(ins)}
    print("function", name, len(ins), "instructions")
    for k, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]:k + 1]
                h = collections.Counter()
                for _, x in body:
                    x = re.sub(r"^@!?U?P\d+\s+", "", x)
                    h[x.split()[0].split(".")[0]] += 1
                if len(body) >= 20:
                    print("loop 0x%x..0x%x: %d instructions: %s" % (tgt, a, len(body),
                          ", ".join("%s %d" % kv for kv in h.most_common())))
