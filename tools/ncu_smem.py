#!/usr/bin/env python
"""Shared-memory wavefronts per load site of one captured kernel (`ncu --page source --csv`): executed count,
wavefronts, ideal wavefronts and average conflict ways, for the hottest LDS/STS instructions.
   python tools/ncu_smem.py gpurun_out/prof_X.ncu-rep [warp_samples]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
nws = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
out = []
tot_w = tot_i = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    w = int(r[ix["L1 Wavefronts Shared"]] or 0)
    if w == 0: continue
    n = int(r[ix["Instructions Executed"]]); ideal = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    out.append((w, n, ideal, r[ix["Source"]].strip()))
    tot_w += w; tot_i += ideal
out.sort(reverse=True)
print(f"shared wavefronts {tot_w} ({tot_w/nws:.1f}/warp-sample), ideal {tot_i} ({tot_i/nws:.1f}/warp-sample)")
for w, n, ideal, src in out[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"  {w/nws:6.2f}/ws  ideal {ideal/nws:5.2f}  x{w/max(n,1):5.2f} per inst ({n/nws:4.2f} inst/ws)  {src[:70]}")
