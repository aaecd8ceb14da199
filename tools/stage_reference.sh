#!/bin/bash
# Stage a copy of the reference's Python sources under baseline/_ref/effocr (git-ignored, travels with gpurun) so that
# tests/test_gpu_reference_drivers.py can run the UNMODIFIED reference scripts on the GPU box, where /root/reference does
# not exist.  Never committed: baseline/_ref/ is in .gitignore.  Usage: bash tools/stage_reference.sh [/root/reference]
set -eu
SRC=${1:-/root/reference}
DST=$(dirname "$0")/../baseline/_ref/effocr
rm -rf "$DST"
mkdir -p "$DST"
cp "$SRC"/infer_effocr.py "$SRC"/infer_effocr_onnx_multi.py "$DST"/
for d in utils models onnx_engines effocr_datasets; do
  mkdir -p "$DST/$d"
  cp "$SRC/$d"/*.py "$DST/$d"/
done
if [ -d "$SRC/english_font_files" ]; then cp -r "$SRC/english_font_files" "$DST"/; fi  # for bench.py --font-dir (2 MB)
echo "staged $(find "$DST" -name '*.py' | wc -l) files under $DST"
