"""Count the Blackwell-specific SASS mnemonics per kernel of libb200vt.so (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce, UBLKCP = 1-D bulk copy,
SYNCS = mbarrier ops, HMMA = mma.sync (the N <= 32 temporal kernel only), UCGABAR = cluster barrier.
    python tools/sass_census.py [lib.so] > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "videotuna-dev_b200", "libb200vt.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA",
        "LDGSTS", "UCGABAR", "MUFU.TANH", "MUFU.EX2", "FFMA2", "FMUL2", "FADD2", "RED.", "ATOM"]
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("vt::(anonymous namespace)::", "").replace("void ", "")
        name = re.sub(r"\((?:const |CUtensorMap|__nv|float|int|long|vt::|unsigned|void\*|bool).*", "", name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p), line):
            counts[name][p] += 1
print(f"# SASS mnemonic census of {os.path.basename(lib)} (sm_100a), kernels with at least one of: {', '.join(pats)}")
for k, c in counts.items():
    if c:
        print(f"{k}\n    " + "  ".join(f"{p}={n}" for p, n in c.items()))
