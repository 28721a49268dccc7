#!/usr/bin/env python
"""Dynamic opcode mix and stall breakdown of one captured kernel from `ncu --page source --csv`.
   python tools/ncu_opmix.py gpurun_out/prof_X.ncu-rep [warp_samples]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
nws = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); stalls = collections.Counter(); samples = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if not m: continue
    n = int(r[ix["Instructions Executed"]])
    ops[m.group(2)] += n
    samples += int(r[ix["# Samples"]])
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]])
tot = sum(ops.values())
print("warp instructions executed:", tot, ("= %.1f per warp-sample" % (tot / nws)) if nws else "")
for k, v in ops.most_common(28):
    print(f"  {k:12s} {v:12d} {100*v/tot:5.1f}%" + (f"  {v/nws:7.1f}/warp-sample" if nws else ""))
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64-pipe (DFMA/DMUL/DADD/DSETP/DMNMX):", fp64, f"{100*fp64/tot:.1f}%", (f"{fp64/nws:.1f}/warp-sample" if nws else ""))
print("stall samples:", samples)
for k, v in stalls.most_common(12):
    print(f"  {k:28s} {v:8d} {100*v/max(samples,1):5.1f}%")
