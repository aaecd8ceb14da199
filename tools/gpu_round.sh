#!/bin/bash
# One gpurun call that refreshes everything the round's evidence needs:
#   GPU parity tests, smoke, both bench arms, the ncu launch list of a bench run, and one
#   `ncu --set full` capture of a ViT-S layer (GEMMs, attention, LayerNorm) at batch 1024.
# Usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log; tail -2 $O/smoke.log
echo "== bench" ; timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 1500 $O/bench.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?"; tail -c 600 $O/bench_reference.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --index-random --pipeline-lines 0 > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
echo "== ncu full (one ViT-S layer)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_tn|attention|layernorm_rows|mlp_fused|proj_ln" -s 5 -c 5 -f -o $O/layer_full \
    python tools/profile_gemm.py 2 > $O/ncu_full.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"crop_resize|knn_" -c 8 -f -o $O/misc_full \
    python tools/profile_misc.py > $O/ncu_misc.log 2>&1; echo "rc=$?"
echo "== attention wait-time timeline"
PYTHONPATH=. timeout 120 python tools/att_timeline.py 0 > $O/att_timeline.txt 2>&1; echo "rc=$?"
ls -la $O
