"""Calibrate the VAE working-set model (bytes per full-resolution pixel-frame) and time a long sliced clip
(BASELINE config 5: latent T=16..32 at 90x160 -> 61..125 frames of 720x1280)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import svr2_import

pkg = svr2_import.load_package()
from comfyui_seedvr2_videoupscaler_b200.vae import B200VideoVAE

sd = pkg.weights.synth_vae_state_dict(seed=1, dtype=torch.float16)
eng = B200VideoVAE(sd)
out = {}
g = torch.Generator().manual_seed(0)


def peak(fn):
    torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    t0 = time.time(); r = fn(); torch.cuda.synchronize()
    return r, torch.cuda.max_memory_allocated() - base, time.time() - t0


for T in (2, 3):
    z = torch.randn(1, 16, T, 90, 160, generator=g).cuda()
    r, pk, dt = peak(lambda: eng.decode(z).sample)
    F = r.shape[2]
    out[f"decode_T{T}"] = {"frames": F, "peak_GB": pk / 1e9, "B_per_pixel_frame": pk / (F * 720 * 1280), "s": dt}
    x = r.float().clamp(-1, 1)
    r2, pk2, dt2 = peak(lambda: eng.encode(x).latent)
    out[f"encode_F{F}"] = {"peak_GB": pk2 / 1e9, "B_per_pixel_frame": pk2 / (F * 720 * 1280), "s": dt2}
    del r, r2, x

for T in (16, 32):
    z = torch.randn(1, 16, T, 90, 160, generator=g).cuda()
    fit = eng._frames_that_fit(720, 1280)
    eng.decode(z[:, :, :2])
    r, pk, dt = peak(lambda: eng.decode(z).sample)
    out[f"long_decode_T{T}"] = {"frames": r.shape[2], "fit_frames": fit, "peak_GB": pk / 1e9, "s": dt,
                                "frames_per_s": r.shape[2] / dt}
    x = r.float().clamp(-1, 1)
    del r
    r2, pk2, dt2 = peak(lambda: eng.encode(x).latent)
    out[f"long_encode_F{x.shape[2]}"] = {"latent": list(r2.shape), "peak_GB": pk2 / 1e9, "s": dt2,
                                         "frames_per_s": x.shape[2] / dt2}
    del x, r2
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/slicing_probe.json", "w"), indent=1)
