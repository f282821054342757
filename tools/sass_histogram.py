#!/usr/bin/env python
"""Opcode histogram of the shipped cubin per kernel (cuobjdump -sass): the mnemonics that show what the kernel is made of --
UBLKCP / SYNCS (TMA bulk copy + mbarrier), BAR (named barriers), LDL / STL (local memory: the contact path's World), MUFU,
ATOMS -- so that round-over-round changes (e.g. local-memory reduction) are visible.  No GPU needed.

    python tools/sass_histogram.py [path/to/lib.so] > profiles/rNN_sass_histogram.json
"""
import collections
import json
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nascargymnasium_b200", "libncg_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WATCH = ("UBLKCP", "SYNCS", "BAR", "LDL", "STL", "LDS", "STS", "LDG", "STG", "MUFU", "ATOMS", "ATOMG", "RED", "FFMA", "FMUL", "FADD", "DFMA", "CALL", "BRA", "BSSY", "WARPSYNC", "SHFL", "HMMA", "UTCHMMA")
out, cur, n = {}, None, 0
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        mm = re.search(r"(ncg_[a-z_]+kernel)(I[\w]*?E)?Ev", name)
        short = name
        if "ncg_step_kernel" in name:
            t = re.search(r"ncg_step_kernelILi(\d)ELi(\d)ELi(\d)ELb(\d)E", name)
            short = "ncg_step_kernel<%s,%s,%s,%s>" % t.groups() + (" [cc unit]" if "ncg_cc" in name or "b200_cc" in name else "") if t else name
        elif mm:
            short = mm.group(1)
        cur = out.setdefault(short, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("BAR", "MUFU") and op.startswith(w)):
                cur[w] += 1
res = {k: {"total": v["total"], "bytes": v["total"] * 16, **{w: v[w] for w in WATCH if v[w]}} for k, v in out.items() if v["total"] > 50}
print(json.dumps({"library": os.path.basename(so), "functions": res}, indent=1))
