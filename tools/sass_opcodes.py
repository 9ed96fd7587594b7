#!/usr/bin/env python3
"""SASS opcode census of libjwavecuda.so (cuobjdump -sass): per kernel the counts of the mnemonics that show which
engines the code uses -- UBLKCP (1-D bulk TMA copy, cp.async.bulk), UBLKPF (bulk L2 prefetch), LDGSTS (cp.async),
SYNCS (mbarrier), DFMA, LDS/STS, BAR -- plus registers.  usage: tools/sass_opcodes.py [regex] > profiles/r2_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

so = "jwave-pro_b200/libjwavecuda.so"
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else ".")
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        counts[cur]["_total"] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts.keys()), capture_output=True, text=True).stdout.splitlines()
keys = ["UBLKCP", "UBLKPF", "UTMALDG", "LDGSTS", "SYNCS", "DFMA", "LDS", "STS", "LDG", "STG", "BAR", "LDCU", "CALL"]
print("# %s: SASS opcode census (sm_100a).  UBLKCP = cp.async.bulk (1-D TMA), UBLKPF = cp.async.bulk.prefetch.L2," % so)
print("# LDGSTS = cp.async, SYNCS = mbarrier; no UTMALDG / UTC*MMA expected: contiguous 1-D rows, fp64 FIR (not a contraction)")
print("# kernel | regs stack | instr | " + " ".join(keys))
tot = collections.Counter()
for (mangled, c), nm in zip(counts.items(), names):
    for k in keys:
        tot[k] += c[k]
    short = re.sub(r"\(anonymous namespace\)::|jwc::|void ", "", nm)
    short = re.sub(r"\(.*", "", short)
    if not pat.search(short):
        continue
    r = regs.get(mangled, (0, 0, 0))
    print("%s | %d %d | %d | %s" % (short, r[0], r[1], c["_total"], " ".join("%s=%d" % (k, c[k]) for k in keys if c[k])))
print("# whole library: " + " ".join("%s=%d" % (k, tot[k]) for k in keys))
