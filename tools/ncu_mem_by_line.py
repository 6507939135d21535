#!/usr/bin/env python
"""Memory traffic of a kernel by CUDA source line, from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
(profile taken with --import-source on; library built with -lineinfo).  Prints, per source line that issues global or
local memory instructions: L1 tag requests, L2 theoretical sectors (global and local), share of the kernel's total, and
the source text.  Usage: ncu_mem_by_line.py <export.csv> [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr, cur_file, out = None, "", []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or r[0] == "" or r[0] == "Function Name":
        continue
    d = dict(zip(hdr, r))  # (the two "Source" columns collide; the CUDA text is r[1])
    def num(k):
        try:
            return float(d.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0
    out.append((cur_file, int(r[0]), r[1].strip(), num("L1 Tag Requests Global"), num("L2 Theoretical Sectors Global"),
                num("L2 Theoretical Sectors Local"), num("Instructions Executed"), num("# Samples"), num("stall_long_sb")))
tot_l1 = sum(o[3] for o in out) or 1.0
tot_l2 = sum(o[4] for o in out) or 1.0
tot_loc = sum(o[5] for o in out) or 1.0
tot_inst = sum(o[6] for o in out) or 1.0
tot_smp = sum(o[7] for o in out) or 1.0
print(f"totals: L1 tag requests (global) {tot_l1:.4g}, L2 theoretical sectors global {tot_l2:.4g}, local {tot_loc:.4g}, "
      f"warp instructions {tot_inst:.4g}, samples {tot_smp:.4g}")
print("  %L2glob   %L1req   %L2local   %inst   %samples  %long_sb_of_samples  file:line  source")
for o in sorted(out, key=lambda o: -(o[4] + o[5]))[:top]:
    if o[4] + o[5] == 0:
        break
    print(f"  {o[4]/tot_l2*100:6.2f}  {o[3]/tot_l1*100:6.2f}  {o[5]/tot_loc*100:8.2f}  {o[6]/tot_inst*100:6.2f}  {o[7]/tot_smp*100:7.2f}  "
          f"{o[8]/tot_smp*100:7.2f}   {o[0]}:{o[1]:<5d} {o[2][:100]}")
