"""Per-kernel SASS opcode counts of the shipped library (static evidence of what each kernel is made of):

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor copy, UBLKCP = cp.async.bulk, HMMA = mma.sync,
FFMA2 / FADD2 / FMUL2 = packed fp32 (sm_100), SYNCS = mbarrier operations."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio-calm_b200", "csrc", "libaudiocalm_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG",
         "SHFL", "SYNCS", "BAR", "LDL", "STL", "DFMA", "DADD"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
        kernels[cur]["_total"] += 1
print("# cuobjdump -sass audio-calm_b200/csrc/libaudiocalm_b200.so: static instruction counts per kernel (sm_100a)")
print("kernel,total," + ",".join(WATCH))
for k, c in kernels.items():
    print(f'"{k}",{c["_total"]},' + ",".join(str(c[w]) for w in WATCH))
