#!/usr/bin/env python
"""Aggregate the per-line table of ncu_by_line.py by function (line ranges found from the sources).
Usage: ncu_by_region.py <lines.txt>"""
import collections
import re
import sys

CSRC = "/root/repo/i3rc_monte_carlo_model_b200/csrc/"


def functions(path):
    out = []
    lines = open(path).read().splitlines()
    for i, ln in enumerate(lines, 1):
        m = re.match(r"(?:I3RC_HD|__global__|static|inline).*?\b(\w+)\(", ln)
        if m and not ln.startswith(" "):
            out.append((i, m.group(1)))
    return out


tabs = {f: functions(CSRC + f) for f in ("transport.cuh", "philox.cuh")}
# kernels.cuh: split k_transport by its phase comments
kl = open(CSRC + "kernels.cuh").read().splitlines()
kt = []
for i, ln in enumerate(kl, 1):
    if "EVENT batch ====" in ln:
        kt.append((i, "K:event batch"))
    elif "TRACE round ====" in ln:
        kt.append((i, "K:trace round (steps)"))
    elif "// close finished rays" in ln:
        kt.append((i, "K:close rays"))
    elif "// idle lanes take the next tasks" in ln:
        kt.append((i, "K:pop tasks"))
    elif "// flush the warp's counters" in ln:
        kt.append((i, "K:epilogue"))
    elif "__global__" in ln and "k_transport" in kl[i] + ln:
        kt.append((i, "K:prologue"))
tabs["kernels.cuh"] = kt


def region(f, l):
    t = tabs.get(f)
    if not t:
        return "other:" + f
    name = "other:" + f
    for start, n in t:
        if start <= l:
            name = n
    return name


agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for ln in open(sys.argv[1]):
    m = re.match(r"\s*([\d.]+)%smp\s+([\d.]+)%inst act=\s*([\d.]+) (\S+?):(\d+)", ln)
    if not m:
        continue
    smp, inst, act, f, l = float(m[1]), float(m[2]), float(m[3]), m[4], int(m[5])
    g = region(f, l)
    agg[g][0] += smp
    agg[g][1] += inst
    agg[g][2] += inst * act
print(open(sys.argv[1]).readline().strip())
for g, (s, i, ia) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{g:32s} smp {s:6.2f}%  inst {i:6.2f}%  act {ia/max(i,1e-9):5.1f}")
