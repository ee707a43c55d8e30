"""Per-source-line stall samples of one kernel from an ncu report.

usage: python profiles/ncu_lines.py report.ncu-rep kernel_regex [file_substring] [top_n]

Runs `ncu -i report --page source --csv --print-source sass,cuda --kernel-name regex:...`, sums `# Samples` of every
SASS instruction onto its CUDA source line (the first launch that matches) and prints the hottest lines with their
dominant stall reasons.
"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per = collections.defaultdict(lambda: [0, collections.Counter(), ""])
cur_file, hdr, seen_fn = "", None, set()
first_kernel_done = False
kernels = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Kernel Name":
        kernels += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or kernels > 1:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    si = hdr.index("# Samples")
    try:
        n = int(r[si])
    except (ValueError, IndexError):
        continue
    if n == 0:
        continue
    key = (cur_file.split("/")[-1], line)
    per[key][0] += n
    per[key][2] = r[1][:100]
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                v = int(r[j])
            except (ValueError, IndexError):
                v = 0
            if v:
                per[key][1][h[6:]] += v
tot = sum(v[0] for v in per.values())
print(f"total samples {tot}")
for (f, line), (n, st, src) in sorted(per.items(), key=lambda x: -x[1][0])[:topn]:
    if fsub and fsub not in f:
        continue
    top = ", ".join(f"{k}:{v}" for k, v in st.most_common(3))
    print(f"{n:7d} {100.0 * n / tot:5.1f}%  {f}:{line:<5d} [{top}]  {src.strip()}")
