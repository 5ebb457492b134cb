"""Clip-level runner over the B200 engines.

Mirrors ``VideoDiffusionInfer`` (reference ``src/core/infer.py``): ``vae_encode``
(:117-199), ``inference`` (:315-395, one Euler step, cfg = 1: x0 = x_t - v,
``samplers/euler.py:59-63``), ``vae_decode`` (:203-278), with the latents handed
between phases on the device (no host bounce).  ``upscale_clip`` strings them
together the way ``generation_phases.py`` does for one clip: 4n+1 temporal pad
(:109-124), clamp + pad-16 + normalise (``generation_utils.py:72-84``), encode,
condition = [latent | 1] (``infer.py:54-78``), DiT, decode, crop, [0,1].
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .dit import B200NaDiT, dit_config
from .vae import B200VideoVAE

SCALING_FACTOR = 0.9152   # configs_3b/main.yaml:60
SHIFTING_FACTOR = 0.0


def pad_4n1(n: int) -> int:
    """frames -> next 4n+1 (generation_phases.py:109-124)."""
    return n if n % 4 == 1 else n + (4 - (n - 1) % 4)


class SeedVR2Engine:
    def __init__(self, dit_cfg: dict, dit_sd: Dict[str, torch.Tensor], vae_sd: Dict[str, torch.Tensor],
                 txt_embed: torch.Tensor, device="cuda"):
        self.device = torch.device(device)
        self.dit = B200NaDiT(dit_cfg, dit_sd, device=device)
        self.vae = B200VideoVAE(vae_sd, device=device)
        self.txt = txt_embed.to(self.device, torch.bfloat16).contiguous()

    # ---- VideoDiffusionInfer.vae_encode ---------------------------------
    @torch.no_grad()
    def vae_encode(self, clip: torch.Tensor) -> torch.Tensor:
        """clip (3,T,H,W) in [-1,1] -> latent (T',h,w,16) bf16, scaled."""
        z = self.vae.encode(clip[None].to(self.device, torch.bfloat16)).latent   # (1,16,T',h,w)
        z = (z - SHIFTING_FACTOR) * SCALING_FACTOR
        return z[0].permute(1, 2, 3, 0).contiguous()

    # ---- VideoDiffusionInfer.inference ------------------------------------
    @torch.no_grad()
    def inference(self, noise: torch.Tensor, latent: torch.Tensor) -> torch.Tensor:
        """noise, latent (T',h,w,16) -> x0 (T',h,w,16).  condition = cat[latent, 1] (task 'sr')."""
        T, h, w, c = latent.shape
        ones = torch.ones(T, h, w, 1, device=self.device, dtype=torch.bfloat16)
        vid = torch.cat([noise.to(self.device, torch.bfloat16), latent.to(torch.bfloat16), ones], -1)
        v = self.dit(vid.view(T * h * w, 2 * c + 1), self.txt, [[T, h, w]], [[self.txt.shape[0]]]).vid_sample
        return noise.to(self.device, torch.bfloat16) - v.view(T, h, w, c)

    # ---- VideoDiffusionInfer.vae_decode -----------------------------------
    @torch.no_grad()
    def vae_decode(self, latent: torch.Tensor) -> torch.Tensor:
        """latent (T',h,w,16) -> sample (3,T,H,W) bf16 in ~[-1,1]."""
        z = latent.permute(3, 0, 1, 2)[None]
        z = z / SCALING_FACTOR + SHIFTING_FACTOR
        return self.vae.decode(z).sample[0]

    # ---- one clip end to end ------------------------------------------------
    @torch.no_grad()
    def upscale_clip(self, frames: torch.Tensor, noise: Optional[torch.Tensor] = None, seed: int = 42) -> torch.Tensor:
        """frames (T,H,W,3) in [0,1], already resized to the target resolution.
        Returns (T,H,W,3) bf16 in [0,1] on the device."""
        T0, H0, W0, _ = frames.shape
        x = frames.to(self.device, torch.bfloat16).clamp(0, 1)
        T = pad_4n1(T0)
        if T > T0:
            x = torch.cat([x, x[-1:].expand(T - T0, -1, -1, -1)], 0)
        ph, pw = (16 - H0 % 16) % 16, (16 - W0 % 16) % 16
        x = x.permute(3, 0, 1, 2)                                  # c t h w
        if ph or pw:
            x = torch.nn.functional.pad(x, (0, pw, 0, ph))
        x = (x - 0.5) / 0.5
        latent = self.vae_encode(x)
        if noise is None:
            g = torch.Generator(device=self.device).manual_seed(seed)
            noise = torch.randn(latent.shape, generator=g, device=self.device, dtype=torch.bfloat16)
        x0 = self.inference(noise, latent)
        y = self.vae_decode(x0)                                     # (3,T,H,W)
        y = y[:, :T0, :H0, :W0].permute(1, 2, 3, 0)
        return (y.float() * 0.5 + 0.5).clamp(0, 1).to(torch.bfloat16)


def build_synthetic_engine(variant="3b", device="cuda", seed=1234, txt_len=58) -> SeedVR2Engine:
    """Random-init weights of the named architecture (no checkpoints exist offline)."""
    from . import weights
    cfg = dit_config(variant)
    dit_sd = weights.synth_dit_state_dict(cfg, seed=seed, dtype=torch.float16, device=device)
    vae_sd = weights.synth_vae_state_dict(seed=seed + 1, dtype=torch.float16, device=device)
    g = torch.Generator().manual_seed(seed + 2)
    txt = torch.randn(txt_len, cfg["txt_in_dim"], generator=g)
    eng = SeedVR2Engine(cfg, dit_sd, vae_sd, txt, device=device)
    del dit_sd, vae_sd
    return eng
