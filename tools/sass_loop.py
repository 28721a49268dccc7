#!/usr/bin/env python
"""Static instruction mix of the streaming kernel's hot loop (the largest backward-branch body) in a built library.
   python tools/sass_loop.py path/to/lib.so [--dump]
Off-GPU proxy for the dynamic mix: the loop body covers two samples per lane (x and y half of a 64-sample group)."""
import re, subprocess, sys, collections
lib = sys.argv[1]
fun = "_ZN4bump13stream_kernelILb0ELb0ELi0EEEvNS_7ColumnsENS_4WorkEPKiPKdPdPy"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = []
for ln in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, t in ins:
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?(\w+)\)?|BRA.*0x([0-9a-f]+)", t)
    m2 = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt < a and (best is None or a - tgt > best[1] - best[0]):
            best = (tgt, a)
# the per-group loop = the LAST backward branch inside the largest loop body (the outer one re-enters after a flush)
lo, hi = best
inner = None
for a, t in ins:
    if "BRA" in t and lo <= a <= hi:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2 and lo < int(m2.group(1), 16) < a and (t.startswith("@") or "BRA.U UP" in t):
            if inner is None or a - int(m2.group(1), 16) > inner[1] - inner[0]:
                inner = (int(m2.group(1), 16), a)
if inner and inner[1] - inner[0] > 0x800:
    lo, hi = inner
# cold regions: bodies skipped by a predicated forward branch of >= 32 instructions (the accumulator rescale, the
# warp flush at an event boundary)
cold = []
for a, t in ins:
    if lo <= a <= hi and t.startswith("@") and "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt > a and tgt - a >= 0x200:
                cold.append((a, tgt))
def is_cold(a):
    return any(c0 < a < c1 for c0, c1 in cold)
body = [(a, t) for a, t in ins if lo <= a <= hi and not is_cold(a)]
print("hot path", hex(lo), hex(hi), "cold regions", [(hex(a), hex(b)) for a, b in cold])
def opname(t):
    t = re.sub(r"^@!?U?P\w+\s+", "", t)
    return t.split()[0].split(".")[0]
cnt = collections.Counter(opname(t) for _, t in body)
fp64 = sum(cnt[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
def nregs(t):
    t2 = re.sub(r"^@!?U?P\w+\s+", "", t)
    ops = t2.split(None, 1)[1] if " " in t2 else ""
    parts = [p.strip() for p in ops.split(",")]
    srcs = parts[1:]
    regs = set()
    for p in srcs:
        m = re.match(r"-?\|?(R\d+)", p)
        if m and not p.startswith("c["):
            if ".reuse" in p:
                continue
            regs.add(m.group(1))
    return len(regs)
three = sum(1 for _, t in body if opname(t) in ("DFMA",) and nregs(t) >= 3)
n = len(body)
print(f"loop [{lo:#x}, {hi:#x}]: {n} instructions for 2 samples -> {n/2:.1f}/sample; FP64-pipe {fp64/2:.1f}/sample; "
      f"other {(n-fp64)/2:.1f}/sample; DFMA with 3 register sources (no reuse) {three/2:.1f}/sample; "
      f"issue model {(2*fp64 + three + (n-fp64))/2:.0f} cycles/sample")
print(" ".join(f"{k}:{v/2:g}" for k, v in cnt.most_common()))
if "--dump" in sys.argv:
    for a, t in body:
        print(f"{a:05x}  {t}")
