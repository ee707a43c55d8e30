"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMEM / TMA mnemonics in every kernel of
libglf_sm100a.so (cuobjdump -sass; runs without a GPU).

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (must be 0), FFMA2 = packed fp32 pairs."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "glfusion_b200", "libglf_sm100a.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "FFMA2", "RED", "SYNCS"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*", "", name).replace("glf::(anonymous namespace)::", "").replace("void ", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        for k in MN:
            if op == k or (k == "RED" and op in ("RED", "REDG")) or (k == "HMMA" and op.startswith("HMMA")):
                cur[k] += 1
print(f"{'kernel':72s} " + " ".join(f"{k:>8s}" for k in MN))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[k] for k in MN):
        continue
    print(f"{name[:72]:72s} " + " ".join(f"{c[k]:8d}" for k in MN))
    tot.update(c)
print(f"{'TOTAL':72s} " + " ".join(f"{tot[k]:8d}" for k in MN))
