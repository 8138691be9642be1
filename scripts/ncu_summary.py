#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics per kernel, and the hottest source lines by stall samples."""
import csv, subprocess, sys, io, collections

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'local_load...']
kern = [r for r in rows[2:] if len(r) > 5]
print("| metric | " + " | ".join(r[idx['Kernel Name']].split('(')[0] for r in kern) + " |")
print("|---|" + "---|" * len(kern))
for w in want:
    if w in idx:
        print(f"| {w} | " + " | ".join(r[idx[w]] for r in kern) + " |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
# split per kernel sections: each begins with a header row containing "Source"
sections = []
cur = None
for line in csv.reader(io.StringIO(src)):
    if not line:
        continue
    if line[0] == "#" or (len(line) > 1 and line[1] == "Source"):
        cur = {"hdr": line, "rows": []}; sections.append(cur); continue
    if cur is not None:
        cur["rows"].append(line)
for si, sec in enumerate(sections):
    h = sec["hdr"]
    try:
        isrc = h.index("Source"); 
        isamp = next(i for i, x in enumerate(h) if x.startswith("# Samples") or x == "Sampling Data (All)" or "Samples" in x)
        iinst = next((i for i, x in enumerate(h) if x.startswith("Instructions Executed")), None)
    except StopIteration:
        print("section", si, "columns:", h[:12]); continue
    tot = sum(float(r[isamp] or 0) for r in sec["rows"] if len(r) > isamp and r[isamp].replace('.', '', 1).isdigit())
    rows2 = sorted((r for r in sec["rows"] if len(r) > isamp and r[isamp].replace('.', '', 1).isdigit()), key=lambda r: -float(r[isamp]))[:top]
    print(f"\n## kernel section {si}: total samples {tot:.0f}")
    for r in rows2:
        extra = f" inst={r[iinst]}" if iinst is not None else ""
        print(f"{float(r[isamp]) / max(tot, 1) * 100:5.1f}%{extra}  L{r[0]}: {r[isrc].strip()[:150]}")
