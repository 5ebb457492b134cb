import os, sys, time, importlib, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
pkg = load_package()
dit = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.dit")
cfg = dit.dit_config("7b")
sd = pkg.weights.synth_dit_state_dict(cfg, seed=7, dtype=torch.float16, device="cuda")
eng = dit.B200NaDiT(cfg, sd); del sd
T, H, W = 2, 270, 480   # config 4 shard: 5 frames -> 2 latent frames at 4K
g = torch.Generator().manual_seed(1)
vid = torch.randn(T * H * W, 33, generator=g).cuda(); txt = torch.randn(58, 5120, generator=g).cuda()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    out = eng(vid, txt, [[T, H, W]], [[58]]).vid_sample
    torch.cuda.synchronize()
    L = T * (H // 2) * (W // 2)
    print(f"7B DiT step L={L}: {(time.time() - t0) * 1e3:.1f} ms ({8.15e9 * L / (time.time() - t0) / 1e12:.0f} TFLOP/s), rms {out.float().pow(2).mean().sqrt():.3f} finite {torch.isfinite(out).all().item()}", flush=True)
