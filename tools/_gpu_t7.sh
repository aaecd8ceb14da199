mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recognizer.py -q -rf -s -k convnext 2>&1 | tail -4
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4b.json 2> gpurun_out/bench_c4b.err; echo "c4 rc=$?"; tail -c 400 gpurun_out/bench_c4b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4b.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')}, d.get('cpu_baseline', {}).get('parity', {}).get('max_rel_embedding_err'))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step'])[:6]: print(' ', k, v)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_launches.csv python tools/profile_c4.py 1024 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/c4_launches.csv')) if len(r) > 10 and r[0].isdigit()]
# second forward only
names = [(r[4].split('(')[0][:60], float(r[-1])) for r in rows]
half = len(names) // 2
for n, t in names[half:]:
    if 'dwconv' in n or 'mlp' in n: print(f'{n:62s} {t/1e3:9.1f} us')
PY
