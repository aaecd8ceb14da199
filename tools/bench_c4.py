"""BASELINE config C4: ConvNeXt-Tiny recognizer, 50k-glyph index, 4096 crops -- crops/s and the kNN stage alone.
Also times ViT-Tiny (config C1's model) for the record."""
import sys
import torch
sys.path.insert(0, ".")
from effocr_b200 import ops, synth
from effocr_b200.encoders import TimmConvNeXtParams, TimmViTParams
from effocr_b200.pipeline import PackedCrops, RecognizerPipeline


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
crops, _ = synth.synthetic_crops(4096, seed=0)
packed = PackedCrops(crops)
px, im, bx, n = packed.to_device()
g = torch.Generator().manual_seed(1)
for name, params, D, nidx, mb in (("convnext_tiny", TimmConvNeXtParams("convnext_tiny"), 768, 50000, 256),
                                  ("vit_tiny_patch16_224", TimmViTParams("vit_tiny_patch16_224"), 192, 94, 1024)):
    sd = {"net." + k: v.detach().clone() for k, v in params.state_dict().items()}
    index = torch.nn.functional.normalize(torch.randn(nidx, D, generator=g), dim=1)
    pipe = RecognizerPipeline(sd, index, max_batch=mb)
    ms = timeit(lambda: pipe.recognize_device(px, im, bx, n, 10))
    emb = pipe.embed_boxes(px, im, bx, n)
    ms_knn = timeit(lambda: pipe.index.search_device(emb, 10), n=10)
    flops = {"convnext_tiny": 8_909_526_528, "vit_tiny_patch16_224": 2_506_982_400}[name]
    print(f"{name}: {n} crops, {nidx}-glyph index, k=10: {ms:.2f} ms -> {n / ms * 1e3:.0f} crops/s "
          f"({flops * n / ms / 1e9:.0f} TFLOP/s encoder); kNN stage alone {ms_knn:.3f} ms "
          f"({2.0 * n * nidx * D / ms_knn / 1e9:.1f} TFLOP/s fp32-equivalent)", flush=True)
    del pipe
