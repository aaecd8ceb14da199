mkdir -p gpurun_out
timeout 300 python tools/mma_numerics_probe.py > gpurun_out/mma_probe.log 2>&1; cat gpurun_out/mma_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step', 'data')})
print('e2e', d['e2e'])
print('parity', d.get('cpu_baseline', {}).get('parity'))
pc = d.get('pipeline_c3') or {}
print({k: pc.get(k) for k in ('lines_per_s', 'crops_per_s', 'localizer_lines_per_s', 'crops_per_line', 'gpu_launches_per_pass', 'parity', 'roofline_localizer')})
PY
