"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass): the Blackwell-specific instructions each kernel
really contains.  Usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cvpr2023-unidistill_b200", "libunidistill_b200.so")
COLS = ["UTMALDG", "UTCHMMA", "UTCQMMA", "UTCBAR", "STTM", "LDTM", "UBLKCP", "LDG.E.EL", "SYNCS", "UTMAPF", "CCTL", "REDG", "RED.E",
        "ATOMG", "F2FP"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|\(.*", "", name)
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for c in COLS:
            if op.startswith(c):
                cur[c] += 1
print("SASS mnemonics per kernel of libunidistill_b200.so (cuobjdump -sass, sm_100a); kernels without any of these columns omitted")
print("%-86s" % "kernel" + "".join("%10s" % c for c in COLS))
for name, c in counts.items():
    if sum(c.values()):
        print("%-86s" % name[:86] + "".join("%10d" % c[k] for k in COLS))
