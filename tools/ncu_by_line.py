#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS) dump by CUDA source line, using `nvdisasm -g` line info of the
cubin the profiled library was built from.  Usage: ncu_by_line.py <sass.csv> <nvdisasm.sass> <kernel mangled substring> [top]"""
import csv
import collections
import re
import sys

csv_path, sass_path, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# --- line table from nvdisasm
lines = {}
cur = None
active = False
for ln in open(sass_path):
    if ln.startswith(".text."):
        active = kern in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        lines[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ia, isrc, ie, it, ismp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = [0, 0, 0]
for r in rows[2:]:
    try:
        addr = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    key = lines.get(addr - base, ("?", 0))
    n, t, s = int(r[ie]), int(r[it]), int(r[ismp])
    a = agg[key]
    a[0] += n
    a[1] += t
    a[2] += s
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[3][hdr[i]] += v
    tot[0] += n
    tot[1] += t
    tot[2] += s
print(f"total warp instructions {tot[0]:.4g}, avg active threads {tot[1]/tot[0]:.2f}, samples {tot[2]}")
src_cache = {}
def src(f, l):
    import glob
    if f not in src_cache:
        c = glob.glob(f"/root/repo/i3rc_monte_carlo_model_b200/csrc/{f}")
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[l - 1].strip()[:90] if 0 < l <= len(L) else ""
for key, (n, t, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    top_st = ",".join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"{s/tot[2]*100:5.2f}%smp {n/tot[0]*100:5.2f}%inst act={t/max(n,1):5.1f} {key[0]}:{key[1]:<4d} {src(*key):90s} {top_st}")
