mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_blocks.py -q -rf -k "ln_gemm" 2>&1 | tail -5
timeout 120 python tools/ab_kernels.py ln qkv ln_qkv
for v in 0 1; do echo "bench EFFOCR_LN_QKV=$v"; EFFOCR_LN_QKV=$v timeout 600 python bench.py --steps 10 --warmup 3 --pipeline-lines 0 > gpurun_out/bench_lnq$v.json 2>gpurun_out/bench_lnq$v.err; tail -c 300 gpurun_out/bench_lnq$v.err; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_lnq$v.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')}, d['cpu_baseline']['parity']['max_rel_embedding_err'], d['cpu_baseline']['parity']['top1_agree'])
    print({k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items() if v['ms_per_step'] > 0.05})
except Exception as e: print('no result', e)
PY
done
