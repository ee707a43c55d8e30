#!/usr/bin/env python
"""Hot SASS of one kernel of an ncu report: python profiles/ncu_hot.py file.ncu-rep KERNEL_ID [top]
Prints opcode histogram weighted by executed instructions and the top stall-sample lines."""
import collections
import csv
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(out) if l.startswith('"Kernel Name"')] + [len(out)]
k = int(kid)
blk = out[starts[k]:starts[k + 1]]
print(blk[0][:150])
rows = list(csv.DictReader(blk[1:]))
hist = collections.Counter()
tot = 0
for r in rows:
    n = int(r["Instructions Executed"] or 0)
    op = r["Source"].split()
    op = [o for o in op if not o.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    hist[name] += n
    tot += n
print("total warp-instructions executed:", tot)
for k, v in hist.most_common(18):
    print(f"  {k:12s} {v:10d} {100.0 * v / tot:5.1f}%")
rows.sort(key=lambda r: -int(r["Warp Stall Sampling (All Samples)"] or 0))
ts = sum(int(r["Warp Stall Sampling (All Samples)"] or 0) for r in rows)
print("top stall lines (of", ts, "samples):")
for r in rows[:top]:
    print(f"  {int(r['Warp Stall Sampling (All Samples)']):6d}  {r['Source'][:100]}")
