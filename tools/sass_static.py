"""Static SASS evidence per kernel of the built library (no GPU needed): instruction counts by mnemonic class, with the
TMA / mbarrier mnemonics (UBLKCP, SYNCS) called out.  python tools/sass_static.py > profiles/rN_sass_static.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "entropy_coders_b200", "libfse_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
print("# static SASS of %s (sm_100a), instructions per kernel; TMA = UBLKCP (cp.async.bulk), mbarrier = SYNCS.*" % os.path.basename(so))
for k, c in per.items():
    tot = sum(c.values())
    fam = collections.Counter()
    for op, n in c.items():
        fam[op.split(".")[0]] += n
    tma = {op: n for op, n in c.items() if op.startswith("UBLKCP") or op.startswith("SYNCS") or op.startswith("UTMA")}
    print("%-40s %6d instr | LDS %4d STS %4d LDG %3d STG %3d SHFL %3d IMAD %4d LOP3 %4d SHF %4d PRMT %3d BAR %2d | TMA/mbarrier: %s"
          % (k[:40], tot, fam["LDS"], fam["STS"], fam["LDG"], fam["STG"], fam["SHFL"], fam["IMAD"], fam["LOP3"], fam["SHF"], fam["PRMT"],
             fam["BAR"], ", ".join("%s x%d" % kv for kv in sorted(tma.items())) or "-"))
