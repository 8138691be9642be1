#!/usr/bin/env python3
"""Print the hot SASS (instructions executed >= frac*max) of one kernel from an .ncu-rep."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
fn = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "Kernel Name": fn = r[1]; hdr = None; continue
    if r[0] == "Address": hdr = r; continue
    if fn and kern in fn and hdr and len(r) > 5: rows.append(r)
h = {k: i for i, k in enumerate(hdr)}
ie, isrc, ismp, ithr = h['Instructions Executed'], h['Source'], h['# Samples'], h['Avg. Threads Executed']
mx = max(float(r[ie]) for r in rows)
hot = [r for r in rows if float(r[ie]) >= frac * mx]
print(f"{kern}: {len(rows)} SASS instructions, {len(hot)} with executed >= {frac}*max ({mx:.3g}); total executed {sum(float(r[ie]) for r in rows):.4g}")
for r in hot:
    st = {k: float(r[h[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}
    top = max(st.items(), key=lambda kv: kv[1])
    print(f"{float(r[ie]):11.0f} thr{float(r[ithr]):5.1f} smp{float(r[ismp]):7.0f} {top[0][6:]:>10s}  {r[isrc][:110]}")
