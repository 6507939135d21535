#!/usr/bin/env python
"""Registers, spills and shared memory of every k_transport instantiation, from the build's ptxas -v log."""
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "i3rc_monte_carlo_model_b200/csrc/_build/api.ptxas.log"
t = open(path).read()
pat = (r"Compiling entry function '(\S+)'[^\n]*\n[^\n]*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
       r"[^\n]*Used (\d+) registers[^\n]*?(\d+) bytes smem")
for m in re.finditer(pat, t):
    n = m.group(1)
    if "k_transport" in n:
        args = re.search(r"k_transportILi(\d+)ELb(\d)ELb(\d)ELb(\d)ELi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)ELb(\d)ELb(\d)ELb(\d)", n)
        print("k_transport<BLOCK=%s REG=%s FAST=%s SPLIT=%s MINB=%s STEPS=%s NSLOT=%s QCAP=%s TSM=%s JUMP=%s TABSM=%s>" % args.groups(),
              "stack", m.group(2), "spill st/ld", m.group(3), m.group(4), "regs", m.group(5), "smem", m.group(6))
