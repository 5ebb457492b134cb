import os, sys, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
load_package()
lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
dev = "cuda"
frames, H, W, C = 2, 2160, 3840, 128
x = torch.randn(frames, H * W, C, device=dev, dtype=torch.bfloat16)
y = torch.empty(2 + frames, H * W, C, device=dev, dtype=torch.bfloat16)
g = torch.ones(C, device=dev, dtype=torch.bfloat16); b = torch.zeros(C, device=dev, dtype=torch.bfloat16)
need = lib.load().svr2_groupnorm_scratch_bytes(frames, H * W, C)
st = torch.empty(need // 8 + 8, device=dev, dtype=torch.float64)
def run():
    lib.call("svr2_groupnorm_bf16", lib.ptr(x), lib.ptr(y), frames, H * W, C, lib.ptr(g), lib.ptr(b), 1e-6, 1, 2, 1, lib.ptr(st), st.numel() * 8, lib.stream())
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"groupnorm {frames}x{H}x{W}x{C}: {ms:.3f} ms, {6.0 * x.numel() / ms / 1e6:.0f} GB/s (6 B/elem)")
