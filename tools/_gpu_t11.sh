mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recognizer.py -q -rf -s -k convnext 2>&1 | tail -4
for v in 0 1; do echo "EFFOCR_CNX_DWCONV=$v"; EFFOCR_CNX_DWCONV=$v timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4d$v.json 2> gpurun_out/bench_c4d$v.err; tail -c 300 gpurun_out/bench_c4d$v.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_c4d$v.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')}, d.get('cpu_baseline', {}).get('parity', {}).get('max_rel_embedding_err'))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:5]: print(' ', k, v)
PY
done
