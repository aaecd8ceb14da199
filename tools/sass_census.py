"""Per-kernel SASS census of libeffocr_b200.so: tcgen05 / TMEM / TMA / legacy tensor-core mnemonics per kernel
(`cuobjdump -sass`), written to profiles/r02_sass_census.txt.  UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, HMMA = mma.sync."""
import re, subprocess, sys, collections
so = sys.argv[1] if len(sys.argv) > 1 else "effocr_b200/libeffocr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
mn = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "HMMA", "MUFU.EX2", "FFMA"]
rows, cur, i = [], None, -1
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = collections.Counter()
        rows.append((names[i], cur))
        continue
    if cur is None:
        continue
    for k in mn:
        if re.search(r"\b" + re.escape(k) + r"\b", line):
            cur[k] += 1
tot = collections.Counter()
print(f"{len(rows)} kernels in {so}")
print(f"{'kernel':78s} " + " ".join(f"{k:>8s}" for k in mn))
for name, c in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r"\(.*", "", name).replace("effocr::", "").replace("void ", "")[:78]
    print(f"{short:78s} " + " ".join(f"{c[k]:8d}" for k in mn))
    tot.update(c)
print(f"{'TOTAL':78s} " + " ".join(f"{tot[k]:8d}" for k in mn))
