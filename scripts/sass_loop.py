#!/usr/bin/env python3
"""Static view of a kernel's SASS: instruction count by opcode inside the outermost loop.

usage: sass_loop.py file.cubin [kernel-substring]
Development aid for counting issue slots per event without a GPU (cuobjdump -sass).
"""
import re
import subprocess
import sys
from collections import Counter


def kernels(cubin):
    text = subprocess.check_output(["cuobjdump", "-sass", cubin], text=True)
    out, name, cur = {}, None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        out[name] = cur
    return out


def main():
    cubin = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ins in kernels(cubin).items():
        if pat not in name:
            continue
        # outermost loop = backward branch with the largest span
        best = None
        for addr, text in ins:
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
        print(f"== {name}: {len(ins)} instructions", end="")
        if not best:
            print(" (no loop)")
            continue
        body = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
        print(f", outer loop {best[0]:#x}..{best[1]:#x}: {len(body)} instructions")
        ops = Counter()
        for _, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t.split()[0].split(".")[0]] += 1
        print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common()))
        if "-v" in sys.argv:
            for a, t in body:
                print(f"    {a:06x}  {t}")


if __name__ == "__main__":
    main()
