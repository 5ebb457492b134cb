"""Launch-bound sizes (BASELINE config 1: one 256^2 -> 512^2 image; a 5-frame 360p clip): eager vs CUDA-graph replay
of the whole clip (pre-process + encode + DiT + decode + format)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svr2_import import load_package
load_package()
import importlib
pipeline = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.pipeline")
eng = pipeline.build_synthetic_engine("3b")
res = {}
for name, (T, h, w, target) in {"cfg1_image_256_to_512": (1, 256, 256, 512), "clip_5f_180p_to_360p": (5, 180, 320, 360),
                                "clip_5f_360p_to_720p": (5, 360, 640, 720)}.items():
    frames = torch.rand(T, h, w, 3, device="cuda", dtype=torch.bfloat16)
    kw = dict(resolution=target)
    def timeit(fn, n=10):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    eager = timeit(lambda: eng.upscale_clip(frames, **kw))
    gc = eng.graphed(frames, **kw)
    graph = timeit(lambda: gc(frames))
    same = torch.equal(gc(frames), eng.upscale_clip(frames, noise=gc.noise, **kw))
    res[name] = {"eager_ms": round(eager, 2), "graph_ms": round(graph, 2), "speedup": round(eager / graph, 2), "identical": bool(same)}
    print(name, res[name], flush=True)
    del gc
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/latency_small.json", "w"), indent=1)
