mkdir -p gpurun_out
for v in 0 1; do echo "EFFOCR_ATT_EXP16=$v"; EFFOCR_ATT_EXP16=$v python tools/ab_kernels.py attention; EFFOCR_ATT_EXP16=$v python - <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from effocr_b200 import ops
torch.manual_seed(0)
batch, heads, T = 8, 6, 197
D = heads * 64
for scale in (1.5, 3.0):
    qkv = (torch.randn(batch * T, 3 * D, device='cuda') * scale).half()
    out = ops.attention(qkv, batch, heads).float()
    q, k, v = qkv.float().reshape(batch, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(batch * T, D)
    print(f'  attention rel err (qkv scale {scale}): {((out - ref).norm() / ref.norm()).item():.2e}')
PY
done
for v in 0 1; do echo "bench EFFOCR_ATT_EXP16=$v"; EFFOCR_ATT_EXP16=$v timeout 600 python bench.py --steps 10 --warmup 3 --pipeline-lines 0 > gpurun_out/bench_att$v.json 2>gpurun_out/bench_att$v.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_att$v.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')}, d['cpu_baseline']['parity']['max_rel_embedding_err'], d['cpu_baseline']['parity']['top1_agree'], d['kernels']['attention'])
PY
done
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"; tail -c 400 gpurun_out/bench_c4.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c4.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tensor_frac_whole_step')}, d['e2e']['value'], d.get('cpu_baseline', {}).get('parity'), d['roofline'])
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']): print(' ', k, v)
PY
timeout 600 python -m pytest tests/test_gpu_transcription.py tests/test_gpu_recognizer.py tests/test_gpu_blocks.py -q -rf -s 2>&1 | grep -v "^\.\+$" | tail -8
