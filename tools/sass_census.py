"""SASS opcode census of libsidekit_b200.so per kernel (tcgen05 / TMEM / bulk-copy / TMA mnemonics of
/opt/skills/guides/B200_PROFILING.md).  Usage: python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sidekit_b200", "libsidekit_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "IMMA", "F2FP",
       "REDUX", "ATOM", "RED", "LDS", "STS", "LDG", "STG", "FFMA", "DFMA", "MUFU")
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in OPS:
            per[cur][op] += 1
tot = collections.Counter()
print("SASS census of %s (sm_100a), %d kernels" % (os.path.basename(lib), len(per)))
print("%-66s %7s  %s" % ("kernel", "instrs", "tensor / TMEM / async-copy and other notable opcodes"))
for k, c in per.items():
    tot.update(c)
    notable = " ".join("%s=%d" % (o, c[o]) for o in OPS if c[o])
    print("%-66s %7d  %s" % (k[:66], c["_total"], notable))
print()
print("totals: " + " ".join("%s=%d" % (o, tot[o]) for o in OPS))
print("tcgen05.mma (UTCHMMA) kernels: %d; TMEM loads (LDTM) kernels: %d; bulk async copies (UBLKCP) kernels: %d; tensor-map TMA "
      "(UTMALDG/UTMASTG) kernels: %d" % (sum(1 for c in per.values() if c["UTCHMMA"]), sum(1 for c in per.values() if c["LDTM"]),
                                        sum(1 for c in per.values() if c["UBLKCP"]), sum(1 for c in per.values() if c["UTMALDG"] or c["UTMASTG"])))
