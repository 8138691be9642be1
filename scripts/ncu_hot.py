#!/usr/bin/env python3
"""Hot source lines of an .ncu-rep (CUDA-C view): samples, instructions executed, avg threads, stall mix."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 15; filt = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = None; fp = None; hdr = None; data = {}
for line in csv.reader(io.StringIO(out)):
    if not line: continue
    if line[0] == "File Path": fp = line[1]; continue
    if line[0] == "Function Name": fn = line[1].split("(")[0]; continue
    if line[0] == "Line No": hdr = line; continue
    if hdr is None or fn is None: continue
    if line[0] == "": continue  # SASS rows; the CUDA-line rows carry the aggregate
    data.setdefault(fn, []).append((fp, hdr, line))
for fn, rows in data.items():
    if filt and filt not in fn: continue
    agg = {}
    for fp, hdr, r in rows:
        h = {k: i for i, k in enumerate(hdr)}
        try:
            samp = float(r[h["# Samples"]]); inst = float(r[h["Instructions Executed"]]); thr = float(r[h["Thread Instructions Executed"]])
        except Exception: continue
        key = (fp.split("/")[-1], r[0], r[1].strip()[:110])
        st = {k: float(r[h[k]] or 0) for k in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_not_selected", "stall_selected", "stall_math", "stall_barrier", "stall_no_inst", "stall_dispatch") if k in h}
        a = agg.setdefault(key, [0, 0, 0, {}]); a[0] += samp; a[1] += inst; a[2] += thr
        for k, v in st.items(): a[3][k] = a[3].get(k, 0) + v
    tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
    print(f"\n### {fn}: samples {tot_s:.0f}, warp instructions {tot_i:.3g}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        stall = ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{a[0]/max(tot_s,1)*100:5.1f}% smp {a[1]/max(tot_i,1)*100:5.1f}% inst thr/inst {a[2]/max(a[1],1):4.1f} | {key[0]}:{key[1]} {key[2]} | {stall}")
