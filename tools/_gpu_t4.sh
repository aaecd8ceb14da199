mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rf -s > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; grep -n "boxes on 32\|passed\|failed\|FAILED" gpurun_out/pytest_gpu3.log | tail
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')})
print('e2e', d['e2e'])
print('parity', d.get('cpu_baseline', {}).get('parity'))
pc = d.get('pipeline_c3') or {}
print({k: pc.get(k) for k in ('lines_per_s', 'crops_per_s', 'localizer_lines_per_s', 'crops_per_line', 'parity')}); print('paths_c5', d.get('paths_c5'))
PY
