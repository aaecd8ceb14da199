"""Turn one `tools/gpu_round.sh` result (gpurun_out/) into the tracked evidence under profiles/.

usage: python tools/collect_profiles.py [round_tag, default r01]
  profiles/<tag>_bench.json, <tag>_bench_reference.json   the two bench lines
  profiles/<tag>_launches_one_step.csv                    the launches of ONE steady-state step from the ncu launch list
  profiles/<tag>_launch_shares.csv                        the same aggregated per kernel
  profiles/<tag>_ncu_layer_summary.txt                    ncu --set full summaries (tools/ncu_summary.py)
  profiles/roofline_traffic.json                          dram read+write bytes per launch per kernel class (bench.py reads it)
"""
import csv
import json
import re
import shutil
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
# optional second argument: output directory (on the GPU box: gpurun_out/profiles, because the .ncu-rep files are too
# large to travel back -- the summaries are made there and the reports deleted)
P = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "profiles"
P.mkdir(parents=True, exist_ok=True)

for src, dst in (("bench.json", f"{tag}_bench.json"), ("bench_reference.json", f"{tag}_bench_reference.json"),
                 ("bench_c4.json", f"{tag}_bench_c4.json"), ("att_timeline.txt", f"{tag}_att_timeline_latest.txt")):
    if (G / src).exists() and (G / src).stat().st_size:
        shutil.copy(G / src, P / dst)


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = name.replace("effocr::", "")
    return re.sub(r"\(.*$", "", name)


# ---- launch list: one steady-state step = from one crop_resize launch to the next
rows = []
with open(G / "launches.csv") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"], r["Block Size"]))
starts = [i for i, r in enumerate(rows) if r[0].startswith("crop_resize_kernel")]
# the 4th step is past the warm-up; steps are identical afterwards
s, e = starts[3], starts[4]
step = rows[s:e]
with open(P / f"{tag}_launches_one_step.csv", "w") as f:
    f.write(f"# one bench step ({len(step)} launches) under `ncu --metrics gpu__time_duration.sum --clock-control none` "
            f"(cold-cache, serialised: compare SHARES); launches {s}..{e - 1} of gpurun_out/launches.csv\n")
    f.write("idx,kernel,grid,block,us\n")
    for i, r in enumerate(step):
        f.write(f'{i},"{r[0]}","{r[2]}","{r[3]}",{r[1]:.3f}\n')
agg = OrderedDict()
for r in step:
    c = agg.setdefault(r[0], [0, 0.0])
    c[0] += 1
    c[1] += r[1]
tot = sum(v[1] for v in agg.values())
with open(P / f"{tag}_launch_shares.csv", "w") as f:
    f.write(f"# one bench step ({len(step)} launches) under ncu duration-only pass (cold-cache, serialised: compare SHARES)\n# total {tot / 1e3:.3f} ms\n")
    f.write("kernel,launches,total_us,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / tot:.3f}\n')

# ---- localizer launch list: the LAST forward pass (profile_yolo.py runs 2 warm-up + 5 timed passes), aggregated per kernel
if (G / "yolo_launches.csv").exists():
    yrows = []
    with open(G / "yolo_launches.csv") as f:
        ylines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(ylines):
        yrows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3))
    ystarts = [i for i, r in enumerate(yrows) if r[0].startswith(("yolo_stem", "yolo_im2col0"))]
    if ystarts:
        last = yrows[ystarts[-1]:]
        yagg = OrderedDict()
        for name, us in last:
            c = yagg.setdefault(name, [0, 0.0])
            c[0] += 1
            c[1] += us
        ytot = sum(v[1] for v in yagg.values())
        with open(P / f"{tag}_yolo_launch_shares.csv", "w") as f:
            f.write(f"# one YOLOv5s forward (64 letterboxed 640x640 lines, split precision; {len(last)} launches) under the ncu duration-only "
                    f"pass (cold-cache, serialised: compare SHARES)\n# total {ytot / 1e3:.3f} ms\n")
            f.write("kernel,launches,total_us,share\n")
            for k, v in sorted(yagg.items(), key=lambda kv: -kv[1][1]):
                f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / ytot:.3f}\n')

# ---- full captures
txt = []
for rep in ("layer_full.ncu-rep", "misc_full.ncu-rep", "yolo_full.ncu-rep"):
    if (G / rep).exists():
        txt.append(f"### {rep}\n" + subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / rep)],
                                                     capture_output=True, text=True).stdout)
(P / f"{tag}_ncu_layer_summary.txt").write_text("\n".join(txt))


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(G / rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    return rr[0], rr[1], rr[2:]


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic_of(rep):
    hdr, units, body = raw(rep)
    k, a, b = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    return [(short(r[k]), float(r[a].replace(",", "")) * SCALE[units[a]] + float(r[b].replace(",", "")) * SCALE[units[b]]) for r in body]


traffic = {}
if (G / "layer_full.ncu-rep").exists():
    # tools/profile_gemm.py: the product kernels (norm1 + QKV, attention, block tail) and the ones they replaced, by name
    names = [("ln_gemm_astat", "gemm_qkv"), ("block_tail", "block_tail"), ("attention", "attention"), ("proj_ln", "proj_ln"),
             ("layernorm_rows", "layernorm"), ("mlp_fused", "mlp_fused"), ("gemm_tn", "gemm_qkv_two_kernel_path")]
    for kname, t in traffic_of("layer_full.ncu-rep"):
        for sub, key in names:
            if sub in kname:
                traffic.setdefault(key, t)
                break
if (G / "misc_full.ncu-rep").exists():
    for kname, t in traffic_of("misc_full.ncu-rep"):
        for key in ("crop_resize", "knn_gemm_topk", "knn_merge_rerank"):
            if kname.startswith(key):
                traffic[key] = t
(P / "roofline_traffic.json").write_text(json.dumps(traffic, indent=1))
print(json.dumps(traffic, indent=1))
print(f"step launches: {len(step)}  total {tot / 1e3:.3f} ms")
