timeout 300 python -m pytest tests/test_gpu_recognizer.py -q -x -k "knn" 2>&1 | tail -2
mkdir -p gpurun_out/knn
timeout 600 python bench.py --steps 20 --warmup 3 --pipeline-lines 0 --paths-lines 0 --no-cpu-baseline > gpurun_out/knn/bench.json 2> gpurun_out/knn/bench.err; echo rc=$?; tail -c 200 gpurun_out/knn/bench.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/knn/bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')})
print({k: round(v['ms_per_step'], 3) for k, v in list(d.get('kernels', {}).items())[:9]})
PY
