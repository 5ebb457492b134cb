import os, sys, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
load_package()
lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
lib.load().svr2_set_cta_pair(1)
torch.manual_seed(0)
for (M, N, K) in [(256, 256, 64), (512, 256, 256), (1000, 768, 320), (4096, 7680, 2560), (300, 128, 128)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    out = lib.linear(a, w)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"pair linear {M}x{N}x{K}: rel err {err:.3e}", flush=True)
