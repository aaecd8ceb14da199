"""BASELINE config C5: the full pipeline over N line FILES sharded across the GPUs of one node.

    python tools/run_c5.py --lines 100000 --out gpurun_out/c5_n1                                      (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_c5.py --lines 100000 --out gpurun_out/c5_n8                                          (8 GPUs, NCCL)

Rank 0 renders `--distinct` synthetic 64 x 1024 lines as PNG files and hard-links them to `--lines` distinct paths (the
decode cost per path is real; the reference's drivers key their results by path).  Every rank builds the engines from the
committed quick-fit weights, the 94-glyph prototype index is embedded on rank 0 with the same kernels and broadcast over
NCCL, `lineio.run_effocr_paths_sharded` transcribes the shards (thread-pool decode one batch ahead of the GPU, no
steady-state collective) and rank 0 gathers the records, writes inference_results.json exactly like the reference's
--save_output block and prints its sha256: the file must be byte-identical for every world size (SURVEY.md section 4,
item 4)."""
import argparse
import hashlib
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=100000)
ap.add_argument("--distinct", type=int, default=2000)
ap.add_argument("--out", default="gpurun_out/c5")
ap.add_argument("--data", default="/tmp/effocr_c5_data")
ap.add_argument("--batch-lines", type=int, default=64)
args = ap.parse_args()

import driver_fixture as DF  # noqa: E402
from effocr_b200 import dist as D, lineio, synth  # noqa: E402
from effocr_b200.infer import EffOCRPipeline  # noqa: E402
from effocr_b200.localizer_engine import EffLocalizer  # noqa: E402
from effocr_b200.pipeline import RecognizerPipeline  # noqa: E402

rank, world, local_rank = D.init_from_env()
torch.cuda.set_device(local_rank)
import torch.distributed as tdist  # noqa: E402


def barrier():
    if world > 1:
        tdist.barrier()
    torch.cuda.synchronize()


# ---- the job: files on disk
data = Path(args.data)
if rank == 0 and not (data / f"done_{args.lines}").exists():
    from PIL import Image

    (data / "src").mkdir(parents=True, exist_ok=True)
    (data / "lines").mkdir(parents=True, exist_ok=True)
    for i, l in enumerate(synth.synthetic_lines(args.distinct, seed=777, tracking=DF.TRACKING)):
        Image.fromarray(l[0]).save(data / "src" / f"s{i:05d}.png")
    for i in range(args.lines):
        dst = data / "lines" / f"line_{i:06d}.png"
        if not dst.exists():
            os.link(data / "src" / f"s{i % args.distinct:05d}.png", dst)
    (data / f"done_{args.lines}").write_text("ok")
barrier()
paths = [str(data / "lines" / f"line_{i:06d}.png") for i in range(args.lines)]

# ---- engines; the index is embedded on rank 0 and broadcast (the only collective besides the final gather)
vsd, ysd = DF.load_npz_state(DF.VIT_WEIGHTS), DF.load_npz_state(DF.YOLO_WEIGHTS)
rec = RecognizerPipeline(vsd, torch.zeros(1, 384), synth.ASCII_GLYPHS, max_batch=2048)
vectors = rec.embed_crops(DF.prototype_crops()).cpu() if rank == 0 else None
vectors, chars = D.broadcast_index(vectors, synth.ASCII_GLYPHS if rank == 0 else None)
rec = RecognizerPipeline(vsd, vectors.cpu(), chars, max_batch=2048)
loc = EffLocalizer(ysd, iou_thresh=0.01, conf_thresh=0.35, input_shape=(640, 640), max_batch=args.batch_lines)
pipe = EffOCRPipeline(loc, rec, chars, lang="en", knn=1)
lineio.run_effocr_paths_sharded(paths[:world * 256], pipe, batch_lines=args.batch_lines)  # warm-up
barrier()
t0 = time.perf_counter()
results, coco = lineio.run_effocr_paths_sharded(paths, pipe, batch_lines=args.batch_lines)
barrier()
dt = time.perf_counter() - t0
if rank == 0:
    os.makedirs(args.out, exist_ok=True)
    keyed = {os.path.basename(k): v for k, v in results.items()}  # the reference re-keys by basename before evaluating
    blob = json.dumps(keyed, indent=2).encode()
    (Path(args.out) / "inference_results.json").write_bytes(blob)
    n_chars = sum(len((v or "").replace(" ", "")) for v in keyed.values())
    summary = {"world_size": world, "lines": args.lines, "distinct_lines": args.distinct, "seconds": dt, "lines_per_s": args.lines / dt,
               "chars_per_s": n_chars / dt, "transcribed": len(keyed), "sha256": hashlib.sha256(blob).hexdigest()}
    (Path(args.out) / "summary.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary), flush=True)
    (Path(args.out) / "inference_results.json").unlink()  # 100k lines: keep the hash and the summary, not 5 MB of JSON
if world > 1:
    tdist.barrier()
    tdist.destroy_process_group()
