#!/bin/bash
# One gpurun call that refreshes everything the round's evidence needs:
#   GPU parity tests, smoke, both bench arms (+ --config c4), the ncu launch lists of a bench run and of the localizer,
#   and `ncu --set full` captures of a ViT-S layer, crop / kNN and the split-precision localizer kernels.
# Usage (from the repo root): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh'; then python tools/collect_profiles.py r02
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q -s -rf > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 300 $O/bench.err
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"
echo "== bench c4" ; timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; echo "rc=$?"
echo "== ncu launch list (bench step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --index-random --pipeline-lines 0 > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu launch list (localizer, split precision)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/yolo_launches.csv \
    python tools/profile_yolo.py 64 > $O/yolo_prof.log 2>&1; echo "rc=$?"
echo "== ncu full (one ViT-S layer)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_tn|attention|layernorm_rows|mlp_fused|proj_ln|ln_gemm_astat|block_tail" -s 8 -c 8 -f -o $O/layer_full \
    python tools/profile_gemm.py 2 > $O/ncu_full.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"crop_resize|knn_" -c 8 -f -o $O/misc_full \
    python tools/profile_misc.py > $O/ncu_misc.log 2>&1; echo "rc=$?"
echo "== ncu full (localizer: the three largest split-precision kernels)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_split3|conv3_tc|yolo_stem_f32" -s 60 -c 12 -f -o $O/yolo_full \
    python tools/profile_yolo.py 64 > $O/ncu_yolo.log 2>&1; echo "rc=$?"
echo "== summaries (the .ncu-rep files are too large to travel back: summarise here, then drop them)"
python tools/collect_profiles.py r02 $O/profiles > $O/collect.log 2>&1; echo "rc=$?"; tail -3 $O/collect.log
for r in layer_full misc_full yolo_full; do
  ncu -i $O/$r.ncu-rep --page source --csv > $O/profiles/r02_${r}_source.csv 2>/dev/null || true
  gzip -f $O/profiles/r02_${r}_source.csv 2>/dev/null || true
done
rm -f $O/*.ncu-rep
du -sh $O; ls -la $O $O/profiles | tail -40
