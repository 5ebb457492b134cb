"""Run on the GPU box: print the PSNRs of the engines vs the reference goldens and the oracle (what the parity tests assert)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT)
from svr2_import import load_package
pkg = load_package()
import importlib
dit = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.dit")
vae = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.vae")
from oracle import dit_oracle, vae_oracle
from oracle.make_golden import DIT_CASES, VAE_CASES, dit_inputs

def psnr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (10 * torch.log10(b.abs().max() ** 2 / (a - b).pow(2).mean())).item()

for name, (variant, over, (T, H, W), l) in DIT_CASES.items():
    cfg = dit.dit_config(variant, **over)
    sd = pkg.weights.synth_dit_state_dict(cfg, seed=1234, dtype=torch.float16)
    vid, txt = dit_inputs(cfg, T, H, W, l)
    gold = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))["out"])
    eng = dit.B200NaDiT(cfg, sd)
    out = eng(vid.cuda(), txt.cuda(), [[T, H, W]], [[l]]).vid_sample
    torch.cuda.synchronize()
    sd32 = {k: v.float() for k, v in sd.items()}
    obf = dit_oracle.dit_forward(sd32, cfg, vid, txt, T, H, W, mode="ref_bf16").float()
    print(f"{name}: engine-vs-golden(fp32 ref) {psnr(out, gold):.1f} dB | oracle_bf16-vs-golden {psnr(obf, gold):.1f} dB | engine-vs-oracle_bf16 {psnr(out, obf):.1f} dB", flush=True)

sdv = pkg.weights.synth_vae_state_dict(seed=4321, dtype=torch.float16)
sdv32 = {k: v.float() for k, v in sdv.items()}
ev = vae.B200VideoVAE(sdv)
for name, (kind, shp) in VAE_CASES.items():
    g = torch.Generator().manual_seed(7)
    gold = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden", name + ".npz"))["out"])
    if kind == "decode":
        T, h, w = shp
        z = torch.randn(1, 16, T, h, w, generator=g)
        out = ev.decode(z.cuda()).sample
        if out.ndim == 4: out = out.unsqueeze(2)
        obf = vae_oracle.vae_decode({k: v.cuda() for k, v in sdv32.items()}, z.cuda(), mode="ref_bf16").float()
    else:
        T, H, W = shp
        x = torch.rand(1, 3, T, H, W, generator=g) * 2 - 1
        out = ev.encode(x.cuda()).latent
        if out.ndim == 4: out = out.unsqueeze(2)
        obf = vae_oracle.vae_encode({k: v.cuda() for k, v in sdv32.items()}, x.cuda(), mode="ref_bf16").float()
    torch.cuda.synchronize()
    print(f"{name}: engine-vs-golden {psnr(out, gold):.1f} dB | oracle_bf16(gpu)-vs-golden {psnr(obf, gold):.1f} dB | engine-vs-oracle_bf16 {psnr(out, obf):.1f} dB", flush=True)

# mid-size timing
for (T, H, W) in [(5, 136, 240)]:
    cfg = dit.dit_config("3b")
    t0 = time.time()
    sd = pkg.weights.synth_dit_state_dict(cfg, seed=1234, dtype=torch.float16, device="cuda")
    eng = dit.B200NaDiT(cfg, sd); del sd
    torch.cuda.synchronize(); print("3B build", time.time() - t0, "s; mem", torch.cuda.memory_allocated() / 2**30, flush=True)
    g = torch.Generator().manual_seed(1)
    vid = torch.randn(T * H * W, 33, generator=g).cuda(); txt = torch.randn(58, 5120, generator=g).cuda()
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        out = eng(vid, txt, [[T, H, W]], [[58]]).vid_sample
        torch.cuda.synchronize(); print(f"3B DiT step {T}x{H}x{W}: {(time.time() - t0) * 1e3:.1f} ms, out rms {out.float().pow(2).mean().sqrt():.3f} finite {torch.isfinite(out).all().item()}", flush=True)
