"""SASS opcode histogram per kernel of libcsparse3_b200.so (cuobjdump -sass), kept under profiles/.

    python tools/sass_hist.py > profiles/sass_histogram_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "csparse3_b200", "libcsparse3_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("csp3::", "")
        kern = re.sub(r"^void ", "", kern)
        kern = re.sub(r"\((?!bool|int).*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# SASS opcode histogram per kernel (sm_100a), %s" % os.path.basename(so))
print("# tensor / TMA mnemonics (UTC*MMA, UTMALDG, DMMA, HMMA) would show here; this path is HBM / shared-memory bound integer and fp64 work")
for k, c in hist.items():
    tot = sum(c.values())
    top = ", ".join("%s %d" % kv for kv in c.most_common(14))
    print("%-60s %6d  %s" % (k[:60], tot, top))
