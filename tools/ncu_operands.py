#!/usr/bin/env python
"""Register-file read pressure of one captured kernel, from `ncu --page source --csv` (SASS view, executed counts).
FP64-pipe instructions are classified by how many 64-bit operands come from the (vector) register file, with and
without counting operands flagged `.reuse` by the previous instruction in the same slot (tools/micro/fp64_mix.cu: a
DFMA with three register operands issues every 3 cycles, with one every 2).
   python tools/ncu_operands.py gpurun_out/prof_X.ncu-rep [warp_samples]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
nws = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
by_regs = collections.Counter(); by_regs_reuse = collections.Counter(); other_regs = collections.Counter()
tot = 0; fp64 = 0
prev_reuse = {}
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]]
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?\s*(.*?)\s*;?\s*$", src)
    if not m: continue
    n = int(r[ix["Instructions Executed"]])
    op = m.group(2)
    ops = [o.strip() for o in m.group(4).split(",")] if m.group(4) else []
    tot += n
    srcs = ops[1:] if op != "DSETP" else ops[2:]      # skip destination(s): DSETP P, PT, a, b, PT
    is_reg = lambda o: re.match(r"^[-|!~]*R\d+", o) is not None
    regs = [o for o in srcs if is_reg(o)]
    fresh = []
    for slot, o in enumerate(srcs):
        if not is_reg(o): continue
        name = re.sub(r"[-|!~]|\.reuse|\.H[01]", "", o).split(".")[0]
        if prev_reuse.get(slot) != name: fresh.append(o)
    prev_reuse = {slot: re.sub(r"[-|!~]|\.reuse", "", o).split(".")[0] for slot, o in enumerate(srcs) if ".reuse" in o}
    if op in FP64:
        fp64 += n
        by_regs[(op, len(regs))] += n
        by_regs_reuse[len(fresh)] += n
    else:
        other_regs[len(regs)] += n
sc = (lambda v: f"{v / nws:8.1f}/warp-sample") if nws else (lambda v: "")
print(f"instructions executed {tot} {sc(tot)}; FP64-pipe {fp64} {sc(fp64)}")
print("FP64-pipe instructions by opcode and number of vector-register source operands:")
for k in sorted(by_regs): print(f"  {k[0]:6s} {k[1]} regs {by_regs[k]:12d} {sc(by_regs[k])}")
print("FP64-pipe instructions by register operands actually read (.reuse hits removed):")
for k in sorted(by_regs_reuse): print(f"  {k} reads {by_regs_reuse[k]:12d} {sc(by_regs_reuse[k])}")
print("other instructions by number of vector-register source operands:")
for k in sorted(other_regs): print(f"  {k} regs {other_regs[k]:12d} {sc(other_regs[k])}")
cyc = sum(max(2, k) * v for k, v in by_regs_reuse.items())
oth = tot - fp64
print(f"model: FP64 issue cycles = sum max(2, reads) = {cyc} {sc(cyc)}; + other instructions x [0.5 .. 1.0] = "
      + (f"{(cyc + 0.5 * oth) / nws:.0f} .. {(cyc + oth) / nws:.0f} cycles/warp-sample" if nws else f"{cyc + 0.5 * oth:.0f} .. {cyc + oth}"))
