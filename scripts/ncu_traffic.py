#!/usr/bin/env python3
"""profiles/rNN_traffic.json from a full ncu capture: DRAM bytes and warp instructions per frame per kernel.
usage: scripts/ncu_traffic.py <rep> <frames in the captured wave> > profiles/r02_traffic.json
bench.py multiplies these per-frame figures by the frames per launch for roofline.traffic and roofline.issue_slots."""
import csv, io, json, subprocess, sys

rep, frames = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def val(r, name):
    return float(r[idx[name]].replace(",", "")) * scale.get(units[idx[name]], 1.0)
out = {"source": "ncu --set full --clock-control none --import-source on, one wave of %d frames x 64 KiB (scripts/r02_ncu.sh)" % frames,
       "frames": frames, "kernels": {}}
for r in rows[2:]:
    if len(r) <= 5:
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("czb::", "")
    if name in out["kernels"]:
        continue  # first launch of each kernel
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out["kernels"][name] = {"dram_read": rd, "dram_write": wr, "bytes_per_frame": (rd + wr) / frames,
                            "warp_inst_per_frame": val(r, "smsp__inst_executed.sum") / frames,
                            "ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]) * {"msecond": 1.0, "usecond": 1e-3, "second": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)}
json.dump(out, sys.stdout, indent=1)
print()
