"""Upsample3D (1x1x1 conv + 3-D pixel shuffle store) at the three decoder shapes of the 4K shard: ms, written GB, TB/s."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svr2_import import load_package
load_package()
lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
dev = "cuda"
for name, (T, H, W, C, temporal) in {"up0 512ch 3x270x480 (t+s)": (3, 270, 480, 512, 1), "up1 512ch 5x540x960 (t+s)": (5, 540, 960, 512, 1),
                                      "up2 256ch 9x1080x1920 (s)": (9, 1080, 1920, 256, 0), "up2 256ch 2x1080x1920 (s)": (2, 1080, 1920, 256, 0)}.items():
    z = 2 if temporal else 1
    x = torch.randn(T, H, W, C, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(4 * z * C, C, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(4 * z * C, device=dev).to(torch.bfloat16)
    T_out = T * z - (1 if temporal else 0)
    y = torch.empty(T_out + 2, 2 * H, 2 * W, C, device=dev, dtype=torch.bfloat16)
    fn = lambda: lib.call("svr2_upsample_shuffle_bf16", lib.ptr(x), T, H, W, C, lib.ptr(w), lib.ptr(b), temporal, 1, lib.ptr(y), 2, 1, lib.stream())
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    gb_w, gb_r = y.numel() * 2 / 1e9, x.numel() * 2 / 1e9
    flops = 2.0 * T * H * W * C * 4 * z * C
    print(f"{name:32s} {ms:7.3f} ms  write {gb_w:6.2f} GB  read {gb_r:5.2f} GB  {(gb_w + gb_r) / ms:6.2f} TB/s  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)
    del x, y
# reference points: pure write (memset / fill) and read+write copy of a buffer of up2's size
y = torch.empty(11, 2160, 3840, 256, device=dev, dtype=torch.bfloat16)
x2 = torch.empty_like(y)
for name, fn, nbytes in (("memset (zero_)", lambda: y.zero_(), y.numel() * 2), ("fill_(1)", lambda: y.fill_(1.0), y.numel() * 2),
                         ("copy_ (read+write)", lambda: y.copy_(x2), 2 * y.numel() * 2)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1):7.3f} ms  {nbytes / 1e9:6.2f} GB  {nbytes / e0.elapsed_time(e1) / 1e9:6.2f} TB/s", flush=True)
