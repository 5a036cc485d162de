#!/usr/bin/env python3
"""SASS opcode histogram of every kernel in liblpcnet_b200.so -> profiles/<tag>_sass_histogram.md
(evidence for which hardware paths the kernels use: IMMA = mma.sync int8 tensor path, UBLKCP = TMA bulk copy,
SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32, LDS/STS/LDG ...).  usage: tools/sass_histogram.py <tag>"""
import collections, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(root, "lpcnet_b200", "liblpcnet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
lines = ["# SASS opcode histogram `%s` (cuobjdump -sass lpcnet_b200/liblpcnet_b200.so)" % tag, "",
         "Base opcode = text before the first '.', full mnemonics of the tensor / TMA / barrier / packed-fp32 instructions listed separately.", ""]
KEY = ("IMMA", "HMMA", "UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "FFMA2", "FMUL2", "FADD2", "IDP", "LDGSTS", "BAR", "MUFU", "LDL", "STL", "REDUX", "SHFL")
for k, c in hist.items():
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    tot = sum(base.values())
    lines += ["## `%s` — %d instructions" % (k, tot), "", "| opcode | count |", "|---|---|"]
    lines += ["| %s | %d |" % (op, n) for op, n in base.most_common(28)]
    special = sorted((op, n) for op, n in c.items() if any(op.startswith(p) for p in KEY))
    if special:
        lines += ["", "notable: " + ", ".join("`%s` x%d" % (op, n) for op, n in special)]
    lines.append("")
out = os.path.join(root, "profiles", "%s_sass_histogram.md" % tag)
open(out, "w").write("\n".join(lines))
print("wrote", out)
