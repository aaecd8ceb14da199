mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_localizer.py tests/test_gpu_e2e_identity.py tests/test_gpu_reference_drivers.py tests/test_gpu_pipeline.py -q -s -rf > gpurun_out/pytest_gpu2.log 2>&1; echo rc=$?
grep -n "yolov5s (\|^\[onnx\]\|^\[torch\]\|boxes on 32\|passed\|failed" gpurun_out/pytest_gpu2.log | tail -30
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import driver_fixture as DF
from effocr_b200.localizer_engine import YoloEngine
ysd = DF.load_npz_state(DF.YOLO_WEIGHTS)
x = torch.rand(64, 3, 640, 640, device='cuda')
for prec in ('split', 'fp16'):
    eng = YoloEngine(ysd, max_batch=64, precision=prec)
    for _ in range(2): eng.forward(x)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): eng.forward(x)
    b.record(); torch.cuda.synchronize()
    print(f'yolo forward 64 x 640x640 {prec}: {a.elapsed_time(b)/5:.3f} ms')
PY
