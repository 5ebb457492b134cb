"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference video VAE
(causal 3-D conv autoencoder, s8_c16_t4_inflation_sd3.yaml).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.

Pinning: ``oracle/make_golden.py`` runs the reference's own
``VideoAutoencoderKLWrapper`` (through ``oracle/ref_import.py``; the only
non-reference arithmetic there is the restated diffusers ``Attention`` forward)
on the same synthetic checkpoint and asserts this restatement matches; outputs
are committed under ``tests/golden/``.

Temporal slicing in the reference (attn_video_vae.py:1254-1300) is exact
(SURVEY.md §4), so this restatement processes the clip un-sliced: every causal
conv sees the first frame replicated ``2*temporal_padding`` times
(causal_inflation_lib.py:232-236, extend_head :423-438).

Layout here is the reference's NCDHW.  ``mode``: "fp32" or "ref_bf16" (all
tensors bf16, as the reference runs the VAE without autocast in bf16).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SCALING_FACTOR = 0.9152  # configs_3b/main.yaml:60
SHIFTING_FACTOR = 0.0


def _dt(mode):
    return torch.bfloat16 if mode == "ref_bf16" else torch.float32


def causal_conv3d(x, w, b, stride=(1, 1, 1), spatial_pad=1, right_pad=False):
    """InflatedCausalConv3d.basic_forward (causal_inflation_lib.py:228-248):
    temporal pad = first frame replicated (kt-1) times [2*temporal_padding for
    kt=3], zero spatial padding, temporal padding 0 in the conv itself.
    right_pad: Downsample3D's (0,1,0,1) asymmetric zero pad with conv padding 0
    (attn_video_vae.py:242-244)."""
    kt = w.shape[2]
    if kt > 1:
        x = torch.cat([x[:, :, :1].expand(-1, -1, kt - 1, -1, -1), x], 2)
    if right_pad:
        x = F.pad(x, (0, 1, 0, 1))
        pad = (0, 0, 0)
    else:
        pad = (0, spatial_pad, spatial_pad)
    return F.conv3d(x, w, b, stride=stride, padding=pad)


def group_norm_per_frame(x, w, b, groups=32, eps=1e-6):
    """causal_norm_wrapper (causal_inflation_lib.py:354-409): GroupNorm on (b t) c h w."""
    n, c, t, h, ww = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(n * t, c, h, ww)
    y = F.group_norm(y, groups, w, b, eps)
    return y.reshape(n, t, c, h, ww).permute(0, 2, 1, 3, 4)


class _P:
    def __init__(self, sd, mode):
        self.sd, self.dt = sd, _dt(mode)

    def __call__(self, k):
        return self.sd[k].to(self.dt)


def resnet_block(P, p, x):
    """ResnetBlock3D.forward (attn_video_vae.py:311-362), temb=None, scale factor 1."""
    h = F.silu(group_norm_per_frame(x, P(p + "norm1.weight"), P(p + "norm1.bias")))
    h = causal_conv3d(h, P(p + "conv1.weight"), P(p + "conv1.bias"))
    h = F.silu(group_norm_per_frame(h, P(p + "norm2.weight"), P(p + "norm2.bias")))
    h = causal_conv3d(h, P(p + "conv2.weight"), P(p + "conv2.bias"))
    if (p + "conv_shortcut.weight") in P.sd:
        x = causal_conv3d(x, P(p + "conv_shortcut.weight"), P(p + "conv_shortcut.bias"), spatial_pad=0)
    return x + h


def mid_attention(P, p, x):
    """UNetMidBlock3D per-frame attention (attn_video_vae.py:656-668) through
    diffusers Attention/AttnProcessor2_0: GN(32) -> q,k,v Linear(+bias) -> 1-head
    SDPA (scale 1/sqrt(C)) -> Linear -> + residual."""
    n, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(n * t, c, h * w)
    y = F.group_norm(xf, 32, P(p + "group_norm.weight"), P(p + "group_norm.bias"), 1e-6).transpose(1, 2)
    q = F.linear(y, P(p + "to_q.weight"), P(p + "to_q.bias"))
    k = F.linear(y, P(p + "to_k.weight"), P(p + "to_k.bias"))
    v = F.linear(y, P(p + "to_v.weight"), P(p + "to_v.bias"))
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, P(p + "to_out.0.weight"), P(p + "to_out.0.bias"))
    o = o.transpose(1, 2) + xf
    return o.reshape(n, t, c, h, w).permute(0, 2, 1, 3, 4)


def mid_block(P, p, x):
    x = resnet_block(P, p + "resnets.0.", x)
    x = mid_attention(P, p + "attentions.0.", x)
    return resnet_block(P, p + "resnets.1.", x)


def upsample3d(P, p, x, temporal: bool):
    """Upsample3D.forward (attn_video_vae.py:110-174): 1x1x1 conv C->rC, pixel
    shuffle 'b (x y z c) f h w -> b c (f z) (h x) (w y)', drop duplicated head
    frame (remove_head, causal_inflation_lib.py:412-419), causal 3x3x3 conv."""
    n, c, f, h, w = x.shape
    z = 2 if temporal else 1
    y = F.conv3d(x, P(p + "upscale_conv.weight"), P(p + "upscale_conv.bias"))
    y = y.view(n, 2, 2, z, c, f, h, w).permute(0, 4, 5, 3, 6, 1, 7, 2).reshape(n, c, f * z, h * 2, w * 2)
    if temporal:
        y = torch.cat([y[:, :, :1], y[:, :, 2:]], 2)
    return causal_conv3d(y, P(p + "conv.weight"), P(p + "conv.bias"))


@torch.no_grad()
def vae_decode(sd: Dict[str, torch.Tensor], z: torch.Tensor, mode: str = "fp32",
               taps: dict | None = None) -> torch.Tensor:
    """Decoder3D.forward (attn_video_vae.py:983-1035).  z (1,16,T,h,w), already
    un-scaled (the runner divides by 0.9152, infer.py:233) -> (1,3,4T-3,8h,8w)."""
    P = _P(sd, mode)
    x = z.to(P.dt)
    x = causal_conv3d(x, P("decoder.conv_in.weight"), P("decoder.conv_in.bias"))
    x = mid_block(P, "decoder.mid_block.", x)
    if taps is not None:
        taps["mid"] = x.float().clone()
    for i in range(4):
        for j in range(3):
            x = resnet_block(P, f"decoder.up_blocks.{i}.resnets.{j}.", x)
        if i < 3:
            x = upsample3d(P, f"decoder.up_blocks.{i}.upsamplers.0.", x, temporal=i < 2)
        if taps is not None:
            taps[f"up{i}"] = x.float().clone()
    x = F.silu(group_norm_per_frame(x, P("decoder.conv_norm_out.weight"), P("decoder.conv_norm_out.bias")))
    return causal_conv3d(x, P("decoder.conv_out.weight"), P("decoder.conv_out.bias"))


@torch.no_grad()
def vae_encode(sd: Dict[str, torch.Tensor], x: torch.Tensor, mode: str = "fp32",
               taps: dict | None = None) -> torch.Tensor:
    """Encoder3D.forward (attn_video_vae.py:808-856) + posterior mode
    (first 16 channels, :1688).  x (1,3,T,H,W) in [-1,1] -> (1,16,(T-1)/4+1,H/8,W/8),
    *not yet* multiplied by the scaling factor."""
    P = _P(sd, mode)
    h = x.to(P.dt)
    h = causal_conv3d(h, P("encoder.conv_in.weight"), P("encoder.conv_in.bias"))
    for i in range(4):
        for j in range(2):
            h = resnet_block(P, f"encoder.down_blocks.{i}.resnets.{j}.", h)
        if i < 3:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv."
            w = P(p + "weight")
            st = 2 if w.shape[2] == 3 else 1
            h = causal_conv3d(h, w, P(p + "bias"), stride=(st, 2, 2), right_pad=True)
        if taps is not None:
            taps[f"down{i}"] = h.float().clone()
    h = mid_block(P, "encoder.mid_block.", h)
    h = F.silu(group_norm_per_frame(h, P("encoder.conv_norm_out.weight"), P("encoder.conv_norm_out.bias")))
    h = causal_conv3d(h, P("encoder.conv_out.weight"), P("encoder.conv_out.bias"))
    return h[:, :16]


def runner_encode(sd, x, mode="fp32"):
    """VideoDiffusionInfer.vae_encode (infer.py:117-199): (mean - shift) * scale,
    returned as (T', h, w, 16)."""
    z = vae_encode(sd, x, mode)
    z = (z - SHIFTING_FACTOR) * SCALING_FACTOR
    return z[0].permute(1, 2, 3, 0)


def runner_decode(sd, lat, mode="fp32"):
    """VideoDiffusionInfer.vae_decode (infer.py:203-278): lat (T',h,w,16) -> z/scale+shift -> decode."""
    z = lat.permute(3, 0, 1, 2)[None]
    z = z / SCALING_FACTOR + SHIFTING_FACTOR
    return vae_decode(sd, z, mode)


def pad_video_temporal(videos: torch.Tensor, count: int = 0, temporal_dim: int = 0, prepend: bool = False) -> torch.Tensor:
    """generation_utils.py:598-657: temporal padding with reversed frames (4n+1 constraint when count == 0)."""
    t = videos.size(temporal_dim)
    if count == 0 and not prepend:
        if t % 4 == 1:
            return videos
        count = ((t - 1) // 4 + 1) * 4 + 1 - t
    if count <= 0:
        return videos
    v = videos.movedim(temporal_dim, 0)
    if count >= t:
        last = v[-1:]
        repeated = last.repeat(count - t + 1, *([1] * (v.ndim - 1)))
        rev = v[1:].flip(0) if t > 1 else last[:0]
        out = torch.cat([repeated, rev, v] if prepend else [v, rev, repeated], 0)
    else:
        rev = v[1:count + 1].flip(0) if prepend else v[-count - 1:-1].flip(0)
        out = torch.cat([rev, v] if prepend else [v, rev], 0)
    return out.movedim(0, temporal_dim)


# --------------------------------------------------------------------------
# spatial tiling (attn_video_vae.py:1302-1630) — optional, off in every BASELINE config; changes results by design
# --------------------------------------------------------------------------
def _tile_plan(extent_h: int, extent_w: int, tile_hw, overlap_hw):
    """Latent-space tile boxes in the reference's visiting order (rows, then columns), skipping a trailing tile that
    lies wholly inside the previous tile's overlap (:1366-1371, :1532-1536)."""
    (th, tw), (oh, ow) = tile_hw, overlap_hw
    sh, sw = max(1, th - oh), max(1, tw - ow)
    boxes = []
    for y0 in range(0, extent_h, sh):
        y1 = min(y0 + th, extent_h)
        for x0 in range(0, extent_w, sw):
            x1 = min(x0 + tw, extent_w)
            if (y0 > 0 and y1 - y0 <= oh) or (x0 > 0 and x1 - x0 <= ow):
                continue
            boxes.append((y0, y1, x0, x1))
    return boxes


def _edge_weights(n: int, ov: int, ramp, fade_lo: bool, fade_hi: bool, like):
    """Separable blend weight of one tile axis: 1 inside, raised-cosine ramp over ``ov`` samples on interior edges only."""
    wgt = torch.ones(n, device=like.device, dtype=like.dtype)
    if ov > 0:
        if fade_lo:
            wgt[:ov] = ramp[:ov]
        if fade_hi:
            wgt[-ov:] = 1 - ramp[:ov]
    return wgt


def _raised_cosine(steps: int, like):
    t = torch.linspace(0, 1, steps=steps, device=like.device, dtype=like.dtype)
    return 0.5 - 0.5 * torch.cos(t * torch.pi)


def _blend_tiles(boxes, tiles, scale: int, ov_hw, ramp_hw, extent_hw, like):
    """Accumulate weighted tiles and normalise by the accumulated weights, in the tiles' dtype with the reference's
    in-place op order (mul_ by the row weights, mul_ by the column weights, +=, addcmul_, div_ by clamp(count))."""
    (H, W), (ovh, ovw) = extent_hw, ov_hw
    first = tiles[0]
    res = torch.zeros(first.shape[0], first.shape[1], first.shape[2], H * scale, W * scale, device=like.device, dtype=first.dtype)
    cnt = torch.zeros(1, 1, 1, H * scale, W * scale, device=like.device, dtype=first.dtype)
    for (y0, y1, x0, x1), tile in zip(boxes, tiles):
        hh = min((y1 - y0) * scale, tile.shape[3], res.shape[3] - y0 * scale)
        ww = min((x1 - x0) * scale, tile.shape[4], res.shape[4] - x0 * scale)
        tile = tile[:, :, : res.shape[2], :hh, :ww].clone()
        wh = _edge_weights(hh, max(0, min(ovh, hh - 1)), ramp_hw[0], y0 > 0, y1 < H, tile).view(1, 1, 1, hh, 1)
        wv = _edge_weights(ww, max(0, min(ovw, ww - 1)), ramp_hw[1], x0 > 0, x1 < W, tile).view(1, 1, 1, 1, ww)
        tile.mul_(wh).mul_(wv)
        ys, xs = y0 * scale, x0 * scale
        res[:, :, : tile.shape[2], ys:ys + hh, xs:xs + ww] += tile
        cnt[:, :, :, ys:ys + hh, xs:xs + ww].addcmul_(wh, wv)
    return res.div_(cnt.clamp(min=1e-6))


def tiled_decode(sd, z: torch.Tensor, tile_size=(512, 512), tile_overlap=(64, 64), mode: str = "fp32", decode_fn=None):
    """VideoAutoencoderKL.tiled_decode (attn_video_vae.py:1472-1630): latent tiles of tile_size // 8, stride = tile -
    overlap // 8, every tile decoded on its own, blended in sample space with ramps of ``tile_overlap`` samples."""
    dec = decode_fn or (lambda t: vae_decode(sd, t, mode))
    _, _, _, H, W = z.shape
    th, tw = max(1, tile_size[0] // 8), max(1, tile_size[1] // 8)
    if H <= th and W <= tw:
        return dec(z)
    loh, low = max(0, min(tile_overlap[0] // 8, th - 1)), max(0, min(tile_overlap[1] // 8, tw - 1))
    boxes = _tile_plan(H, W, (th, tw), (loh, low))
    zt = z.to(_dt(mode)) if decode_fn is None else z
    ramps = (_raised_cosine(tile_overlap[0], zt) if tile_overlap[0] > 0 else None,
             _raised_cosine(tile_overlap[1], zt) if tile_overlap[1] > 0 else None)
    tiles = [dec(z[:, :, :, y0:y1, x0:x1]) for (y0, y1, x0, x1) in boxes]
    return _blend_tiles(boxes, tiles, 8, tile_overlap, ramps, (H, W), zt)


def tiled_encode(sd, x: torch.Tensor, tile_size=(512, 512), tile_overlap=(64, 64), mode: str = "fp32", encode_fn=None):
    """VideoAutoencoderKL.tiled_encode (attn_video_vae.py:1302-1470): sample-space crops of whole latent tiles, every
    crop encoded on its own, blended in latent space with ramps of ``tile_overlap // 8`` latent samples."""
    enc = encode_fn or (lambda t: vae_encode(sd, t, mode))
    _, _, _, H, W = x.shape
    if H <= tile_size[0] and W <= tile_size[1]:
        return enc(x)
    th, tw = max(1, tile_size[0] // 8), max(1, tile_size[1] // 8)
    loh, low = max(0, min(tile_overlap[0] // 8, th - 1)), max(0, min(tile_overlap[1] // 8, tw - 1))
    Hl, Wl = (H + 7) // 8, (W + 7) // 8
    boxes = _tile_plan(Hl, Wl, (th, tw), (loh, low))
    xt = x.to(_dt(mode)) if encode_fn is None else x
    ramps = (_raised_cosine(loh, xt) if loh > 0 else None, _raised_cosine(low, xt) if low > 0 else None)
    tiles = [enc(x[:, :, :, y0 * 8:min(y1 * 8, H), x0 * 8:min(x1 * 8, W)]) for (y0, y1, x0, x1) in boxes]
    return _blend_tiles(boxes, tiles, 1, (loh, low), ramps, (Hl, Wl), xt)
