import os, sys, torch, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
pkg = load_package()
vae = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.vae")
sd = pkg.weights.synth_vae_state_dict(seed=4321, dtype=torch.float16)
eng = vae.B200VideoVAE(sd)
log = []
for name in ("_conv", "_gn", "_attention", "_upsample"):
    orig = getattr(eng, name)
    def wrap(*a, _orig=orig, _name=name, **k):
        out = _orig(*a, **k)
        log.append((_name, a[1] if len(a) > 1 else "", out.body.clone(), out.T))
        return out
    setattr(eng, name, wrap)
g = torch.Generator().manual_seed(3)
z = torch.randn(1, 16, 4, 10, 16, generator=g).cuda()
eng.decode(z); full = list(log); log.clear()
eng.decode(z[:, :, :2]); pre = list(log); log.clear()
eng.decode(z[:, :, :2]); pre2 = list(log); log.clear()
for (n1, p1, a, Ta), (n2, p2, b, Tb), (_, _, c, _) in zip(full, pre, pre2):
    d = (a[:Tb].float() - b.float()).abs().max().item()
    d2 = (c.float() - b.float()).abs().max().item()
    print(f"{n1:10s} {str(p1)[-40:]:40s} T {Ta}/{Tb} max|d| prefix {d:.4g}  rerun {d2:.4g}  |max| {b.float().abs().max().item():.3g}")
