"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Makes the *reference's own* model code (``/root/reference/src/models/...``)
importable in the build container, where three of its third-party
dependencies are absent (``rotary_embedding_torch``, ``diffusers``,
``omegaconf``).  The stubs below restate only the published semantics of the
symbols the reference touches; every line of
model math that runs afterwards is the reference's.

Used by ``oracle/make_golden.py`` to (1) validate the restatements in
``oracle/dit_oracle.py`` / ``oracle/vae_oracle.py`` and (2) generate the
fixtures committed under ``tests/golden/``.  ``/root/reference`` does not
exist on the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or
``bench.py`` imports this file.

Third-party packages restated (unpinned by the reference,
``requirements.txt:8-11``):
  * rotary_embedding_torch >= 0.5.3  (RotaryEmbedding, apply_rotary_emb)
  * diffusers >= 0.33.1              (timestep embedding, Attention forward,
                                      2D block constructors, VAE output types)
"""
from __future__ import annotations

import math
import sys
import types
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn

REFERENCE_ROOT = "/root/reference"


# --------------------------------------------------------------------------
# rotary_embedding_torch (lucidrains) — semantics used at
# dit_3b/rope.py:28-32,46,77-81,118-126 and dit_7b/rope.py
# --------------------------------------------------------------------------
def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], x.shape[-1] // 2, 2)
    x1, x2 = x.unbind(-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rotary_emb(freqs, t, start_index=0, scale=1.0, seq_dim=-2, freqs_seq_dim=None):
    dtype = t.dtype
    if freqs_seq_dim is None and (freqs.ndim == 2 or t.ndim == 3):
        freqs_seq_dim = 0
    if t.ndim == 3 or freqs_seq_dim is not None:
        seq_len = t.shape[seq_dim]
        idx = [slice(None)] * freqs.ndim
        idx[freqs_seq_dim] = slice(-seq_len, None)
        freqs = freqs[tuple(idx)]
    rot_dim = freqs.shape[-1]
    end_index = start_index + rot_dim
    t_left, t_mid, t_right = t[..., :start_index], t[..., start_index:end_index], t[..., end_index:]
    t_mid = (t_mid * freqs.cos() * scale) + (_rotate_half(t_mid) * freqs.sin() * scale)
    return torch.cat((t_left, t_mid, t_right), dim=-1).type(dtype)


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, custom_freqs=None, freqs_for="lang", theta=10000, max_freq=10,
                 num_freqs=1, learned_freq=False, **_):
        super().__init__()
        if freqs_for == "lang":
            freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))
        elif freqs_for == "pixel":
            freqs = torch.linspace(1.0, max_freq / 2, dim // 2) * math.pi
        else:
            raise ValueError(freqs_for)
        self.freqs_for = freqs_for
        self.freqs = nn.Parameter(freqs, requires_grad=learned_freq)

    def forward(self, t, seq_len=None, offset=0):
        freqs = self.freqs
        freqs = torch.einsum("..., f -> ... f", t.type(freqs.dtype), freqs)
        return torch.repeat_interleave(freqs, 2, dim=-1)

    def get_axial_freqs(self, *dims):
        all_freqs = []
        for ind, dim in enumerate(dims):
            if self.freqs_for == "pixel":
                pos = torch.linspace(-1, 1, steps=dim, device=self.freqs.device)
            else:
                pos = torch.arange(dim, device=self.freqs.device)
            freqs = self.forward(pos, seq_len=dim)
            shape = [1] * len(dims) + [freqs.shape[-1]]
            shape[ind] = dim
            all_freqs.append(freqs.reshape(shape))
        all_freqs = torch.broadcast_tensors(*all_freqs)
        return torch.cat(all_freqs, dim=-1)


# --------------------------------------------------------------------------
# diffusers
# --------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False,
                           downscale_freq_shift=1, scale=1, max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class _DiffusersRMSNorm(nn.Module):  # import-only in shipped configs
    def __init__(self, dim, eps, elementwise_affine=True, bias=False):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0,
                 temb_channels=512, groups=32, groups_out=None, pre_norm=True, eps=1e-6,
                 non_linearity="swish", skip_time_act=False, time_embedding_norm="default",
                 kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False, down=False,
                 conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.up, self.down = up, down
        self.output_scale_factor = output_scale_factor
        self.time_embedding_norm = time_embedding_norm
        self.skip_time_act = skip_time_act
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = None
        self.norm2 = nn.GroupNorm(groups_out, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_2d_out_channels = conv_2d_out_channels or out_channels
        self.conv2 = nn.Conv2d(out_channels, conv_2d_out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = self.downsample = None
        self.use_in_shortcut = (in_channels != conv_2d_out_channels
                                if use_in_shortcut is None else use_in_shortcut)
        self.conv_shortcut = None
        if self.use_in_shortcut:
            self.conv_shortcut = nn.Conv2d(in_channels, conv_2d_out_channels, 1, bias=conv_shortcut_bias)


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None,
                 name="conv", kernel_size=None, padding=1, norm_type=None, eps=None,
                 elementwise_affine=None, bias=True, interpolate=True):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_conv_transpose = use_conv_transpose
        self.name = name
        self.interpolate = interpolate
        self.norm = None
        conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding, bias=bias) if use_conv else None
        if name == "conv":
            self.conv = conv
        else:
            self.Conv2d_0 = conv


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv",
                 kernel_size=3, norm_type=None, eps=None, elementwise_affine=None, bias=True):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.padding = padding
        self.name = name
        self.norm = None
        if use_conv:
            conv = nn.Conv2d(channels, self.out_channels, kernel_size, stride=2, padding=padding, bias=bias)
        else:
            conv = nn.AvgPool2d(kernel_size=2, stride=2)
        if name == "conv":
            self.Conv2d_0 = conv
            self.conv = conv
        elif name == "Conv2d_0":
            self.conv = conv
        else:
            self.conv = conv


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32,
                 resnet_pre_norm=True, output_scale_factor=1.0, add_downsample=True,
                 downsample_padding=1):
        super().__init__()  # reference replaces resnets / downsamplers entirely


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, resolution_idx=None, dropout=0.0, num_layers=1,
                 resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish",
                 resnet_groups=32, resnet_pre_norm=True, output_scale_factor=1.0,
                 add_upsample=True, temb_channels=None):
        super().__init__()


class LoRACompatibleConv(nn.Conv2d):
    pass


class SpatialNorm(nn.Module):
    pass


class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention with AttnProcessor2_0, as
    configured at attn_video_vae.py:612-632 (1 head, group norm, residual)."""

    def __init__(self, query_dim, heads=8, dim_head=64, rescale_output_factor=1.0, eps=1e-5,
                 norm_num_groups=None, spatial_norm_dim=None, residual_connection=False,
                 bias=False, upcast_softmax=False, _from_deprecated_attn_block=False, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True)
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, temb=None, **_):
        residual = hidden_states
        b, c, h, w = hidden_states.shape
        x = hidden_states.view(b, c, h * w).transpose(1, 2)
        x = self.group_norm(x.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        hd = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        x = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        x = x.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        x = self.to_out[0](x)
        x = self.to_out[1](x)
        x = x.transpose(-1, -2).reshape(b, c, h, w)
        if self.residual_connection:
            x = x + residual
        return x / self.rescale_output_factor


@dataclass
class DecoderOutput:
    sample: torch.Tensor


@dataclass
class AutoencoderKLOutput:
    latent_dist: "DiagonalGaussianDistribution"


class DiagonalGaussianDistribution:
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.use_slicing = False

    def enable_slicing(self):
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    @property
    def device(self):
        return next(self.parameters()).device


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install_stubs():
    global _installed
    if _installed:
        return
    _installed = True
    _module("rotary_embedding_torch", RotaryEmbedding=RotaryEmbedding, apply_rotary_emb=apply_rotary_emb)
    d = _module("diffusers", AutoencoderKL=AutoencoderKL)
    d.__path__ = []
    models = _module("diffusers.models")
    models.__path__ = []
    _module("diffusers.models.embeddings", get_timestep_embedding=get_timestep_embedding)
    _module("diffusers.models.normalization", RMSNorm=_DiffusersRMSNorm)
    _module("diffusers.models.attention_processor", Attention=Attention, SpatialNorm=SpatialNorm)
    ae = _module("diffusers.models.autoencoders")
    ae.__path__ = []
    _module("diffusers.models.autoencoders.vae", DecoderOutput=DecoderOutput,
            DiagonalGaussianDistribution=DiagonalGaussianDistribution)
    _module("diffusers.models.downsampling", Downsample2D=Downsample2D)
    _module("diffusers.models.lora", LoRACompatibleConv=LoRACompatibleConv)
    _module("diffusers.models.modeling_outputs", AutoencoderKLOutput=AutoencoderKLOutput)
    _module("diffusers.models.resnet", ResnetBlock2D=ResnetBlock2D)
    un = _module("diffusers.models.unets")
    un.__path__ = []
    _module("diffusers.models.unets.unet_2d_blocks", DownEncoderBlock2D=DownEncoderBlock2D,
            UpDecoderBlock2D=UpDecoderBlock2D)
    _module("diffusers.models.upsampling", Upsample2D=Upsample2D)
    u = _module("diffusers.utils", is_torch_version=lambda *a: True)
    u.__path__ = []
    _module("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference_dit(variant="3b"):
    install_stubs()
    import importlib
    return importlib.import_module(f"src.models.dit_{variant}.nadit")


def import_reference_vae():
    install_stubs()
    import importlib
    return importlib.import_module("src.models.video_vae_v3.modules.attn_video_vae")
