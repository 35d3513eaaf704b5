#!/usr/bin/env python3
"""Static view of the ensemble pass: the SASS instructions on the hot path of the inner pass loop (development aid).

usage: sass_path.py file.{o,cubin} kernel_name [-v]       e.g.  sass_path.py build/sys_vilar.o rb_ssa_sys_Vilar_dyn
The pass loop is the innermost loop that holds both MUFU.RCP64H (the divide) and I2F.F64.U64 (the uniform).  The walk
follows fall-through and takes a forward conditional branch when the region it skips is a side exit (it contains a
CALL, a global load/store or a BREAK), i.e. it counts what a warp executes when no lane needs the ziggurat's slow
path and no lane crosses a grid point.  No GPU needed (cuobjdump -sass).
"""
import re, subprocess, sys
from collections import Counter
cubin, pat = sys.argv[1], sys.argv[2]
text = subprocess.check_output(["cuobjdump", "-sass", "-fun", pat, cubin], text=True)
ins = []
for line in text.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
# inner loop: backward BRA.U with smallest span containing >= 10 DSETP
best = None
for a, t in ins:
    m = re.search(r"\bBRA(\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(2), 16)
        if tgt < a:
            n = min(sum(1 for b, u in ins if tgt <= b <= a and "MUFU.RCP64H" in u), sum(1 for b, u in ins if tgt <= b <= a and "I2F.F64.U64" in u))
            if n >= 1 and (best is None or a - tgt < best[1] - best[0]): best = (tgt, a)
lo, hi = best
def region_has_call(a0, a1):
    return any(("CALL" in t) for a, t in ins if a0 <= a < a1)
pc = lo; path = []; seen_first = False
steps = 0
while pc <= hi and steps < 5000:
    steps += 1
    i = addr2i[pc]; a, t = ins[i]
    path.append((a, t))
    m = re.search(r"\bBRA(\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(2), 16)
        cond = t.startswith("@") or "UP" in t.split("BRA")[1].split(",")[0] if "BRA.U" in t else t.startswith("@")
        if tgt <= a: break
        if not cond: pc = tgt; continue
        # conditional forward: take it if the skipped region contains a CALL or rare ops (STG/LDG/I2F second) ; never take if target beyond loop end - 0x40 (liveness skip)
        skipped = [u for b, u in ins if a < b < tgt]
        rare = any(("CALL" in u or "STG" in u or "LDG" in u or "BREAK" in u) for u in skipped)
        if tgt >= hi - 0x30 and not seen_first and a - lo < 0x80: seen_first = True; pc = ins[i + 1][0]; continue
        if rare: pc = tgt; continue
    pc = ins[i + 1][0]
ops = Counter()
for _, t in path:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t.split()[0].split(".")[0]] += 1
print(f"{pat}: inner loop {lo:#x}..{hi:#x}, hot path {len(path)} instructions")
print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common()))
if "-v" in sys.argv:
    for a, t in path: print(f"   {a:05x} {t}")
