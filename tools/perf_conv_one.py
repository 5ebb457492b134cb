"""One representative launch of each dominant kernel for ncu captures (round-1 profiles)."""
import os, sys, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
load_package()
lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
dev = "cuda"
which = sys.argv[1]
def rnd(*s): return torch.randn(*s, device=dev, dtype=torch.bfloat16)
if which == "conv256":      # pair kernel (cta_group::2), 256 -> 256 at 2 x 1080 x 1920
    T, H, W, C = 2, 1080, 1920, 256
    x = rnd(T + 2, H, W, C); w = rnd(C, 27 * C) * 0.01; b = rnd(C); y = torch.empty(T, H, W, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.conv3d(x, T + 2, H, W, C, w, C, (3, 3, 3), 1, 1, 1, T, y, bias=b)
elif which == "conv128":    # swap-AB kernel, 128 -> 128 at 2 x 2160 x 3840
    T, H, W, C = 2, 2160, 3840, 128
    x = rnd(T + 2, H, W, C); w = rnd(C, 27 * C) * 0.01; b = rnd(C); y = torch.empty(T, H, W, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.conv3d(x, T + 2, H, W, C, w, C, (3, 3, 3), 1, 1, 1, T, y, bias=b)
elif which == "conv_sc":    # ResnetBlock3D conv2 + fused 1x1x1 shortcut, 128 (+256) -> 128 at 2 x 2160 x 3840
    import ctypes
    T, H, W, C, C2 = 2, 2160, 3840, 128, 256
    x = rnd(T + 2, H, W, C); x2 = rnd(T, H, W, C2); w = rnd(C, 27 * C + C2) * 0.01; b = rnd(C)
    y = torch.empty(T, H, W, C, device=dev, dtype=torch.bfloat16)
    slots = ctypes.c_int(lib.load().svr2_conv_stat_slots(C, H, W))
    part = torch.empty(T * slots.value * (C // 8) * 4, device=dev, dtype=torch.float32)
    fn = lambda: lib.call("svr2_conv3d_shortcut_stats_bf16", lib.ptr(x), T + 2, H, W, C, lib.ptr(w), C, 3, 3, 3, T, lib.ptr(b),
                          lib.ptr(x2), C2, lib.ptr(y), 0, 0, lib.ptr(part), part.numel() * 4, ctypes.byref(slots), lib.stream())
elif which == "swiglu":     # DiT SwiGLU input projection at the 4K shard (L = 97200)
    L = 97200
    a = rnd(L, 2560); w = rnd(13824, 2560) * 0.02
    fn = lambda: lib.linear(a, w, epi=lib.EPI_SWIGLU)
elif which == "attn":       # DiT window attention at the 4K shard: 243 windows of 405 + 58 tokens, 20 heads
    lens = [463] * 243; total = sum(lens)
    q, k, v = rnd(total, 20, 128), rnd(total, 20, 128), rnd(total, 20, 128)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    o = torch.empty_like(q)
    fn = lambda: lib.attn_varlen(q, k, v, cu, 463, out=o)
elif which == "upsample":   # Upsample3D 1x1x1 conv + pixel shuffle, 256 ch at 2 x 1080 x 1920 -> 2 x 2160 x 3840
    T, H, W, C = 2, 1080, 1920, 256
    x = rnd(T, H, W, C); w = rnd(4 * C, C) * 0.05; b = rnd(4 * C)
    y = torch.empty(T + 2, 2 * H, 2 * W, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.call("svr2_upsample_shuffle_bf16", lib.ptr(x), T, H, W, C, lib.ptr(w), lib.ptr(b), 0, 1,
                          lib.ptr(y), 2, 1, lib.stream())
elif which == "shortcut":   # 1x1x1 conv_shortcut 256 -> 128 at 2 x 2160 x 3840 (swap-AB, 4 k-blocks per tile)
    T, H, W, C = 2, 2160, 3840, 256
    x = rnd(T, H, W, C); w = rnd(128, C) * 0.05; b = rnd(128); y = torch.empty(T, H, W, 128, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.conv3d(x, T, H, W, C, w, 128, (1, 1, 1), 1, 1, 0, T, y, bias=b)
for _ in range(3): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("PERF_REPS", "1"))
e0.record()
for _ in range(n): fn()
e1.record(); torch.cuda.synchronize()
print(which, f"{e0.elapsed_time(e1) / n:.3f} ms")
