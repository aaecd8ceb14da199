"""Top sampled SASS instructions of a kernel in an .ncu-rep (source page): where the warps stall."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= si or not r[si].isdigit():
        continue
    data.append((int(r[si]), r[so].strip(), r[ie], i))
tot = sum(d[0] for d in data) or 1
print(rows[0][1][:100] if rows and len(rows[0]) > 1 else "", "total samples", tot)
for s, src, ie_, i in sorted(data, reverse=True)[:int(sys.argv[4]) if len(sys.argv) > 4 else 22]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{i:5d} exec={ie_:>9}  {src[:100]}")
