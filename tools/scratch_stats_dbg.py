import os, sys, importlib, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from svr2_import import load_package
load_package()
lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
if os.environ.get("SVR2_AB_LIB"):
    lib.LIB_PATH = os.path.abspath(os.environ["SVR2_AB_LIB"])
from test_ops_gpu import rnd, bf, _to_ndhwc
DEV = "cuda"
def run(Cin, Cout, T, H, W, keep):
    x = rnd(1, Cin, T, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, 3, std=(Cin * 27) ** -0.5, seed=2)
    b = rnd(Cout, seed=3)
    x_nd = _to_ndhwc(x, 2)
    w_k = bf(w.permute(0, 2, 3, 4, 1).reshape(Cout, -1)).contiguous()
    y = torch.zeros(T, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    b16 = bf(b) if keep else None
    args = (lib.ptr(x_nd), T + 2, H, W, Cin, lib.ptr(w_k), Cout, 3, 3, 3, 1, 1, 1, T, lib.EPI_BIAS,
            lib.ptr(b16 if keep else bf(b)), None, lib.ptr(y), 0, 0, Cout)
    slots = ctypes.c_int(0)
    assert lib.load().svr2_conv3d_stats_bf16(*args, None, 0, ctypes.byref(slots), lib.stream()) == 0
    part = torch.full((T * slots.value * (Cout // 8) * 4,), float("nan"), device=DEV)
    print("  bias ptr", hex(args[15].value), "part ptr", hex(part.data_ptr()), "part bytes", part.numel() * 4)
    lib.call("svr2_conv3d_stats_bf16", *args, lib.ptr(part), part.numel() * 4, ctypes.byref(slots), lib.stream())
    torch.cuda.synchronize()
    bad = ~torch.isfinite(part.view(T, slots.value, Cout // 8, 4))
    print(Cin, Cout, T, H, W, "keep" if keep else "temp", "slots", slots.value, "bad", int(bad.sum()), "of", bad.numel(),
          "y nan", int((~torch.isfinite(y.float())).sum()))
for keep in (False, True):
    for c in ((64, 128, 2, 20, 36), (128, 256, 2, 19, 30), (256, 512, 1, 12, 20)):
        run(*c, keep)
