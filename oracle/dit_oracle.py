"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference NaDiT forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module; the product path
(``comfyui_seedvr2_videoupscaler_b200``) never does.

Pinning: ``oracle/make_golden.py`` runs the *reference's own* ``NaDiT``
(imported through ``oracle/ref_import.py``) on the same synthetic checkpoints
and asserts this restatement matches it (fp32, max |d| ~1e-5); the outputs are
committed under ``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py``.

Every function cites the reference file:line (relative to /root/reference) it
follows.  ``mode``:
  * "fp32"     — everything in float32 (the numerical ground truth);
  * "ref_bf16" — mirrors the rounding points of the reference's CUDA path
                 (fp16 checkpoint, ``torch.autocast(bf16)``).
"""
from __future__ import annotations

import math
from math import ceil
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

TXT_ROPE_NOTE = "text position (i,i,i); video position (t+l, h, w) — dit_3b/rope.py:172-173"


# --------------------------------------------------------------------------
# configs (configs_3b/main.yaml:6-37, configs_7b/main.yaml:6-33)
# --------------------------------------------------------------------------
def dit_config(variant: str = "3b", **over) -> dict:
    if variant == "3b":
        cfg = dict(variant="3b", dim=2560, heads=20, head_dim=128, layers=32, mm_layers=10,
                   mlp="swiglu", txt_in_dim=5120, in_ch=33, out_ch=16, eps=1e-5,
                   out_norm=True, rope="mm_lang", last_vid_only=True)
    elif variant == "7b":
        cfg = dict(variant="7b", dim=3072, heads=24, head_dim=128, layers=36, mm_layers=36,
                   mlp="gelu", txt_in_dim=5120, in_ch=33, out_ch=16, eps=1e-5,
                   out_norm=False, rope="vid_pixel", last_vid_only=False)
    else:
        raise ValueError(variant)
    cfg.update(over)
    return cfg


def swiglu_hidden(dim: int, expand_ratio: int = 4, multiple_of: int = 256) -> int:
    """dit_3b/mlp.py:54-55"""
    h = int(2 * dim * expand_ratio / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


# --------------------------------------------------------------------------
# windows (dit_3b/window.py:28-83) — boxes in reference enumeration order
# --------------------------------------------------------------------------
def window_boxes(t: int, h: int, w: int, shifted: bool, num_windows=(4, 3, 3)) -> List[Tuple[int, ...]]:
    rnt, rnh, rnw = num_windows
    scale = math.sqrt((45 * 80) / (h * w))
    rh, rw = round(h * scale), round(w * scale)
    wh, ww = ceil(rh / rnh), ceil(rw / rnw)
    wt = ceil(min(t, 30) / rnt)
    out = []
    if not shifted:  # window.py:28-49
        nt, nh, nw = ceil(t / wt), ceil(h / wh), ceil(w / ww)
        for iw in range(nw):
            w0, w1 = iw * ww, min((iw + 1) * ww, w)
            if w1 <= w0:
                continue
            for ih in range(nh):
                h0, h1 = ih * wh, min((ih + 1) * wh, h)
                if h1 <= h0:
                    continue
                for it in range(nt):
                    t0, t1 = it * wt, min((it + 1) * wt, t)
                    if t1 <= t0:
                        continue
                    out.append((t0, t1, h0, h1, w0, w1))
        return out
    # window.py:51-83
    st, sh, sw = (0.5 if wt < t else 0, 0.5 if wh < h else 0, 0.5 if ww < w else 0)
    nt, nh, nw = ceil((t - st) / wt), ceil((h - sh) / wh), ceil((w - sw) / ww)
    nt, nh, nw = (nt + 1 if st > 0 else 1, nh + 1 if sh > 0 else 1, nw + 1 if sw > 0 else 1)
    for iw in range(nw):
        w0, w1 = max(int((iw - sw) * ww), 0), min(int((iw - sw + 1) * ww), w)
        if w1 <= w0:
            continue
        for ih in range(nh):
            h0, h1 = max(int((ih - sh) * wh), 0), min(int((ih - sh + 1) * wh), h)
            if h1 <= h0:
                continue
            for it in range(nt):
                t0, t1 = max(int((it - st) * wt), 0), min(int((it - st + 1) * wt), t)
                if t1 <= t0:
                    continue
                out.append((t0, t1, h0, h1, w0, w1))
    return out


def window_token_index(t: int, h: int, w: int, boxes) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Gather permutation of na.window_idx (dit_3b/na.py:618-641): tokens of each
    box in (t,h,w) row-major order, boxes concatenated.  Returns (tgt_idx,
    window lengths, local (t,h,w) coordinates + box dims per token)."""
    grid = torch.arange(t * h * w).view(t, h, w)
    idx, lens, loc = [], [], []
    for (t0, t1, h0, h1, w0, w1) in boxes:
        sub = grid[t0:t1, h0:h1, w0:w1].reshape(-1)
        idx.append(sub)
        lens.append(sub.numel())
        tt, hh, ww_ = torch.meshgrid(torch.arange(t1 - t0), torch.arange(h1 - h0),
                                     torch.arange(w1 - w0), indexing="ij")
        dims = torch.tensor([t1 - t0, h1 - h0, w1 - w0]).expand(sub.numel(), 3)
        loc.append(torch.cat([torch.stack([tt, hh, ww_], -1).reshape(-1, 3), dims], -1))
    return torch.cat(idx), torch.tensor(lens), torch.cat(loc)


# --------------------------------------------------------------------------
# RoPE tables (rotary_embedding_torch semantics)
# --------------------------------------------------------------------------
def rope_angles_lang(freqs: torch.Tensor, npos: int) -> torch.Tensor:
    """angle[p, j] = p * freqs[j] computed in the buffer's dtype
    (RotaryEmbedding.forward: ``t.type(freqs.dtype)``; G4)."""
    pos = torch.arange(npos)
    return torch.einsum("p,f->pf", pos.type(freqs.dtype), freqs)


def rope_cos_sin_3b(freqs: torch.Tensor, loc: torch.Tensor, txt_len: int):
    """dit_3b/rope.py:130-176.  Returns per-window-token (cos, sin) of shape
    (Lwin, 126) and text (txt_len, 126); cos/sin evaluated in the buffer dtype
    (``freqs.cos()`` inside apply_rotary_emb) then promoted to fp32."""
    nf = freqs.numel()
    npos = int(max(loc[:, 0].max() + txt_len, loc[:, 1].max(), loc[:, 2].max(), txt_len)) + 1
    ang = rope_angles_lang(freqs, npos)                      # (npos, nf)
    ang = torch.repeat_interleave(ang, 2, dim=-1)            # (npos, 2nf) pairs share an angle
    vid = torch.cat([ang[loc[:, 0] + txt_len], ang[loc[:, 1]], ang[loc[:, 2]]], -1)
    ti = torch.arange(txt_len)
    txt = torch.cat([ang[ti]] * 3, -1)
    return (vid.cos().float(), vid.sin().float()), (txt.cos().float(), txt.sin().float())


def rope_cos_sin_7b(freqs: torch.Tensor, loc: torch.Tensor):
    """dit_7b/rope.py:73-111: pixel freqs, positions linspace(-1,1,size) on every
    axis of the *window* shape; video only."""
    outs = []
    for ax in range(3):
        n = loc[:, 3 + ax]
        p = loc[:, ax]
        # linspace(-1, 1, n)[p]; n == 1 -> -1
        cache: Dict[int, torch.Tensor] = {}
        pos = torch.empty(loc.shape[0], dtype=freqs.dtype)
        for nn_ in n.unique().tolist():
            cache[nn_] = torch.linspace(-1, 1, steps=nn_).type(freqs.dtype)
            m = n == nn_
            pos[m] = cache[nn_][p[m]]
        ang = torch.einsum("p,f->pf", pos, freqs)
        outs.append(torch.repeat_interleave(ang, 2, dim=-1))
    vid = torch.cat(outs, -1)
    return vid.cos().float(), vid.sin().float()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """apply_rotary_emb on the first cos.shape[-1] dims; x (L, heads, d) fp32."""
    r = cos.shape[-1]
    xm = x[..., :r]
    x1, x2 = xm[..., 0::2], xm[..., 1::2]
    rot = torch.stack((-x2, x1), -1).reshape(xm.shape)
    xm = xm * cos[:, None, :] + rot * sin[:, None, :]
    return torch.cat([xm, x[..., r:]], -1)


# --------------------------------------------------------------------------
# helpers with reference rounding points
# --------------------------------------------------------------------------
class _Flow:
    def __init__(self, mode: str):
        assert mode in ("fp32", "ref_bf16")
        self.bf = mode == "ref_bf16"

    def act(self, x):  # activation stream dtype
        return x.bfloat16() if self.bf else x.float()

    def lin(self, x, w, b=None):
        """nn.Linear under autocast: operands cast to bf16, fp32 accumulate, bf16 out."""
        if self.bf:
            return F.linear(x.bfloat16(), w.bfloat16(), None if b is None else b.bfloat16())
        return F.linear(x.float(), w.float(), None if b is None else b.float())

    def rnd(self, x):  # a bf16 rounding point
        return x.bfloat16() if self.bf else x


def rms(x: torch.Tensor, eps: float) -> torch.Tensor:
    """dit_3b/normalization.py:88-109 — fp32 under autocast (pow is an fp32-list op)."""
    xf = x.float()
    return xf / torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + eps)


def time_embedding(sd, fl: _Flow, t: float) -> torch.Tensor:
    """dit_3b/embedding.py:25-62 + diffusers get_timestep_embedding(256, flip=False, shift=0)."""
    half = 128
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = torch.tensor([t], dtype=torch.float32)[:, None] * f[None]
    e = torch.cat([a.sin(), a.cos()], -1).to(sd["emb_in.proj_in.weight"].device)
    e = fl.act(e)
    e = fl.lin(e, sd["emb_in.proj_in.weight"], sd["emb_in.proj_in.bias"])
    e = F.silu(e)
    e = fl.lin(e, sd["emb_in.proj_hid.weight"], sd["emb_in.proj_hid.bias"])
    e = F.silu(e)
    e = fl.lin(e, sd["emb_in.proj_out.weight"], sd["emb_in.proj_out.bias"])
    return e  # (1, 6d)


def modulation_vectors(sd, cfg, emb: torch.Tensor) -> Dict[str, torch.Tensor]:
    """All input-independent AdaSingle vectors, fp32:  E[:, layer, g] + P
    (dit_3b/modulation.py:76,100-116).  emb is bf16 in the reference flow and
    the learned vectors are checkpoint dtype, so the sum promotes to fp32."""
    d = cfg["dim"]
    E = emb.float().view(d, 2, 3)  # (d, layer{attn,mlp}, {shift,scale,gate})
    out = {}
    for i in range(cfg["layers"]):
        shared = i >= cfg["mm_layers"]
        last = cfg["last_vid_only"] and i == cfg["layers"] - 1
        for s in ("vid", "txt"):
            if last and s == "txt":
                continue
            key = "all" if shared else s
            for li, layer in enumerate(("attn", "mlp")):
                for gi, g in enumerate(("shift", "scale", "gate")):
                    p = sd[f"blocks.{i}.ada.{key}.{layer}_{g}"].float()
                    out[f"{i}.{s}.{layer}_{g}"] = E[:, li, gi] + p
    if cfg["out_norm"]:
        # G1: vid_out_ada reuses the *attn* slice of emb (nadit.py:236-244 + modulation.py:80-81)
        out["out_shift"] = E[:, 0, 0] + sd["vid_out_ada.out_shift"].float()
        out["out_scale"] = E[:, 0, 1] + sd["vid_out_ada.out_scale"].float()
    return out


def varlen_attention(q, k, v, lens: List[int], bf: bool, impl: str = "math") -> torch.Tensor:
    """dit_3b/attention.py:27-64: per-sequence SDPA, non-causal, scale 1/sqrt(d).
    ``impl`` (ref_bf16 flow on a GPU only): "math" = explicit fp32 scores, bf16 probabilities (the rounding flow of a
    fused kernel, runs anywhere); "sdpa" = torch's own fused bf16 SDPA per window, the reference's default backend
    (attention.py:52-60); "flash_attn" = flash_attn_varlen_func (compatibility.py:321-330).  The latter two give the
    reference-vs-reference noise floor."""
    if bf and impl == "flash_attn":
        from flash_attn import flash_attn_varlen_func
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=q.device)
        return flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens))
    outs, o = [], 0
    for n in lens:
        qi, ki, vi = (x[o:o + n].permute(1, 0, 2).unsqueeze(0) for x in (q, k, v))
        if bf and impl == "sdpa":
            oi = F.scaled_dot_product_attention(qi, ki, vi)
        elif bf:
            # fused kernels: fp32 scores/softmax, probabilities rounded to bf16 for P·V
            s = (qi.float() @ ki.float().transpose(-1, -2)) / math.sqrt(q.shape[-1])
            p = torch.softmax(s, -1)
            oi = (p.bfloat16().float() @ vi.float()).bfloat16()
        else:
            oi = F.scaled_dot_product_attention(qi, ki, vi)
        outs.append(oi.squeeze(0).permute(1, 0, 2))
        o += n
    return torch.cat(outs, 0)


# --------------------------------------------------------------------------
# the forward
# --------------------------------------------------------------------------
@torch.no_grad()
def dit_forward(sd: Dict[str, torch.Tensor], cfg: dict, vid: torch.Tensor, txt: torch.Tensor,
                T: int, H: int, W: int, timestep: float = 1000.0, mode: str = "fp32",
                taps: dict | None = None, attn_impl: str = "math") -> torch.Tensor:
    """NaDiT.forward (dit_3b/nadit.py:190-248, dit_7b/nadit.py:152-190), b = 1.

    vid (T*H*W, 33) latent-pixel rows; txt (l, 5120).  Returns vid_sample (T*H*W, 16).
    """
    fl = _Flow(mode)
    dev = vid.device     # weights and inputs may live on a GPU (parity tests at benchmark sizes); index/RoPE tables are
    #                      built on the host exactly as before and moved over
    d, nh, hd, eps = cfg["dim"], cfg["heads"], cfg["head_dim"], cfg["eps"]
    is7 = cfg["variant"] == "7b"
    vid, txt = fl.act(vid), fl.act(txt)
    l = txt.shape[0]

    # stem -------------------------------------------------------------
    txt = fl.lin(txt, sd["txt_in.weight"], sd["txt_in.bias"])                       # nadit.py:211
    Hp, Wp = H // 2, W // 2
    x = vid.view(T, Hp, 2, Wp, 2, -1).permute(0, 1, 3, 2, 4, 5).reshape(T * Hp * Wp, -1)  # patch_v1.py:91
    x = fl.lin(x, sd["vid_in.proj.weight"], sd["vid_in.proj.bias"])                  # patch_v1.py:96
    emb = time_embedding(sd, fl, timestep)
    mod = modulation_vectors(sd, cfg, emb)
    if taps is not None:
        taps["emb"] = emb.float().clone()
        taps["vid_in"] = x.float().clone()
        taps["txt_in"] = txt.float().clone()

    layouts = []
    for shifted in (False, True):
        boxes = window_boxes(T, Hp, Wp, shifted)
        tgt, lens, loc = window_token_index(T, Hp, Wp, boxes)
        layouts.append((tgt.to(dev), lens.tolist(), loc, torch.argsort(tgt).to(dev)))
    rope_cache: Dict[tuple, tuple] = {}

    for i in range(cfg["layers"]):
        shared = i >= cfg["mm_layers"]
        last = cfg["last_vid_only"] and i == cfg["layers"] - 1
        kv, kt = ("all", "all") if shared else ("vid", "txt")
        pre = f"blocks.{i}."

        # ---- attention branch (mmsr_block.py:107-114) -----------------
        a_v = rms(x, eps) * mod[f"{i}.vid.attn_scale"] + mod[f"{i}.vid.attn_shift"]
        a_t = rms(txt, eps)
        if not last:
            a_t = a_t * mod[f"{i}.txt.attn_scale"] + mod[f"{i}.txt.attn_shift"]
        qkv_v = fl.lin(a_v, sd[pre + f"attn.proj_qkv.{kv}.weight"])                  # mmattn.py:173
        qkv_t = fl.lin(a_t, sd[pre + f"attn.proj_qkv.{kt}.weight"])
        tgt, lens, loc, src = layouts[i % 2]
        qkv_v = qkv_v[tgt].view(-1, 3, nh, hd)                                       # mmattn.py:199-201
        qkv_t = qkv_t.view(-1, 3, nh, hd)
        q_v, k_v, v_v = qkv_v.unbind(1)
        q_t, k_t, v_t = qkv_t.unbind(1)
        q_v = rms(q_v, eps) * sd[pre + f"attn.norm_q.{kv}.weight"].float()           # mmattn.py:207-208
        k_v = rms(k_v, eps) * sd[pre + f"attn.norm_k.{kv}.weight"].float()
        q_t = rms(q_t, eps) * sd[pre + f"attn.norm_q.{kt}.weight"].float()
        k_t = rms(k_t, eps) * sd[pre + f"attn.norm_k.{kt}.weight"].float()
        freqs = sd[pre + "attn.rope.rope.freqs"].cpu()
        rk = (i % 2, freqs.dtype, tuple(freqs.tolist()))                             # same table for every layer that
        if rk not in rope_cache:                                                     # shares layout and frequencies
            if not is7:
                rope_cache[rk] = tuple(t.to(dev) for pair in rope_cos_sin_3b(freqs, loc, l) for t in pair)  # rope.py:130-176
            else:
                rope_cache[rk] = tuple(t.to(dev) for t in rope_cos_sin_7b(freqs, loc))
        if not is7:
            cv, sv, ct, st = rope_cache[rk]
            q_v, k_v = apply_rope(q_v, cv, sv), apply_rope(k_v, cv, sv)
            q_t, k_t = apply_rope(q_t, ct, st), apply_rope(k_t, ct, st)
        else:
            cv, sv = rope_cache[rk]
            q_v, k_v = apply_rope(q_v, cv, sv), apply_rope(k_v, cv, sv)
        # concat text after every window (na.py:320-424) and run varlen attention
        qs, ks, vs, o = [], [], [], 0
        for n in lens:
            qs += [q_v[o:o + n], q_t]
            ks += [k_v[o:o + n], k_t]
            vs += [v_v[o:o + n].float(), v_t.float()]
            o += n
        q_all, k_all, v_all = (fl.act(torch.cat(z)) for z in (qs, ks, vs))           # attention.py:118-121
        out = varlen_attention(q_all, k_all, v_all, [n + l for n in lens], fl.bf, attn_impl).float()  # mmattn.py:257
        ov, ot, o = [], [], 0
        for n in lens:
            ov.append(out[o:o + n])
            ot.append(out[o + n:o + n + l])
            o += n + l
        o_v = torch.cat(ov)[src].reshape(-1, nh * hd)                                # window_reverse mmattn.py:264
        o_t = torch.stack(ot, 0).mean(0).reshape(-1, nh * hd)                        # na.py:396-417
        o_v = fl.lin(o_v, sd[pre + f"attn.proj_out.{kv}.weight"], sd[pre + f"attn.proj_out.{kv}.bias"])
        o_t = fl.lin(o_t, sd[pre + f"attn.proj_out.{kt}.weight"], sd[pre + f"attn.proj_out.{kt}.bias"])
        o_v = fl.rnd(o_v * mod[f"{i}.vid.attn_gate"])                                # modulation.py:112 (in-place on bf16)
        if not last:
            o_t = fl.rnd(o_t * mod[f"{i}.txt.attn_gate"])
        h_v = fl.rnd(o_v + x)
        h_t = fl.rnd(o_t + txt)

        # ---- MLP branch (mmsr_block.py:116-126) ------------------------
        def mlp(hh, key, s):
            m = fl.rnd(rms(hh, eps))                                                  # cast back :119-122
            m = fl.rnd(fl.rnd(m * mod[f"{i}.{s}.mlp_scale"]) + mod[f"{i}.{s}.mlp_shift"])
            if cfg["mlp"] == "swiglu":                                                # mlp.py:60-62
                g = fl.lin(m, sd[pre + f"mlp.{key}.proj_in_gate.weight"])
                u = fl.lin(m, sd[pre + f"mlp.{key}.proj_in.weight"])
                z = fl.rnd(F.silu(g) * u)
                y = fl.lin(z, sd[pre + f"mlp.{key}.proj_out.weight"])
            else:                                                                      # dit_7b/mlp.py:35-43
                u = fl.lin(m, sd[pre + f"mlp.{key}.proj_in.weight"], sd[pre + f"mlp.{key}.proj_in.bias"])
                z = F.gelu(u, approximate="tanh")
                y = fl.lin(z, sd[pre + f"mlp.{key}.proj_out.weight"], sd[pre + f"mlp.{key}.proj_out.bias"])
            y = fl.rnd(y * mod[f"{i}.{s}.mlp_gate"])
            return fl.rnd(y + hh)

        x = mlp(h_v, kv, "vid")
        txt = fl.rnd(h_t + h_t) if last else mlp(h_t, kt, "txt")                     # vid_only: txt_mlp = txt_attn
        if taps is not None:
            taps[f"block{i}"] = x.float().clone()

    # head (nadit.py:234-247) -------------------------------------------
    if cfg["out_norm"]:
        v = rms(x, eps) * sd["vid_out_norm.weight"].float()
        v = v * mod["out_scale"] + mod["out_shift"]
    else:
        v = x
    v = fl.lin(v, sd["vid_out.proj.weight"], sd["vid_out.proj.bias"])                # patch_v1.py:114
    v = v.view(T, Hp, Wp, 2, 2, -1).permute(0, 1, 3, 2, 4, 5).reshape(T * H * W, -1)  # patch_v1.py:121
    return v


def one_step_latent(noise: torch.Tensor, v_pred: torch.Tensor) -> torch.Tensor:
    """EulerSampler with steps=1, v_lerp schedule at t=T: x0 = x_t - v
    (samplers/euler.py:59-63, schedules/base.py:108-110, schedules/lerp.py:44-48)."""
    return noise - v_pred
