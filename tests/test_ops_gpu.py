"""-m gpu: every C-ABI op of libsvr2.so against a plain torch fp32 restatement of the same
reference op (floating point kernels; tolerances are stated per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(DEV)


def bf(x):
    return x.to(torch.bfloat16)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert math.isfinite(e) and e < tol, f"{what}: rel err {e:.3e} >= {tol}"


# ------------------------------------------------------------------ linear
@pytest.mark.parametrize("M,N,K", [(300, 256, 192), (1000, 768, 256), (77, 64, 2560), (513, 384, 320),
                                   (2048, 7680, 2560), (129, 16, 128), (4, 2560, 256), (1, 1536, 256)])
def test_linear_plain(svr2lib, M, N, K):
    a, w = bf(rnd(M, K, seed=1)), bf(rnd(N, K, std=K ** -0.5, seed=2))
    out = svr2lib.linear(a, w)
    ref = a.float() @ w.float().T
    assert_close(out, ref, 4e-3, f"linear {M}x{N}x{K}")


def test_linear_epilogues(svr2lib):
    M, N, K = 777, 512, 448
    a, w = bf(rnd(M, K, seed=1)), bf(rnd(N, K, std=K ** -0.5, seed=2))
    bias, gate = bf(rnd(N, seed=3)), rnd(N, seed=4).float()
    res = bf(rnd(M, N, seed=5))
    acc = a.float() @ w.float().T
    # bias + gate + residual with the reference's rounding points
    out = svr2lib.linear(a, w, bias=bias, gate=gate, residual=res)
    t = bf(acc + bias.float()).float()
    t = bf(t * gate).float()
    ref = bf(t + res.float())
    assert_close(out, ref, 3e-3, "bias+gate+residual")
    out = svr2lib.linear(a, w, bias=bias, epi=svr2lib.EPI_GELU)
    assert_close(out, F.gelu(bf(acc + bias.float()).float(), approximate="tanh"), 4e-3, "gelu")
    out = svr2lib.linear(a, w, bias=bias, epi=svr2lib.EPI_SILU)
    assert_close(out, F.silu(bf(acc + bias.float()).float()), 4e-3, "silu")
    out = svr2lib.linear(a, w, epi=svr2lib.EPI_F32, out_scale=0.25)
    assert out.dtype == torch.float32
    assert_close(out, acc * 0.25, 1e-5, "f32")
    # SwiGLU: tile j of 256 weight rows = [128 gate rows ; 128 in rows]
    g, u = acc[:, :256], acc[:, 256:]
    wg, wu = w[:256], w[256:]
    w_il = torch.cat([wg[:128], wu[:128], wg[128:], wu[128:]], 0).contiguous()
    out = svr2lib.linear(a, w_il, epi=svr2lib.EPI_SWIGLU)
    ref = bf(F.silu(bf(g).float())).float() * bf(u).float()
    assert out.shape == (M, 256)
    assert_close(out, ref, 4e-3, "swiglu")


def test_linear_strided_and_big_k(svr2lib):
    M, N, K = 640, 256, 6912
    a_full = bf(rnd(M, K + 64, seed=7))
    a = a_full[:, :K]
    w = bf(rnd(N, K, std=K ** -0.5, seed=8))
    out = svr2lib.linear(a, w)
    assert_close(out, a.float() @ w.float().T, 4e-3, "strided A")


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("lens,heads", [([1, 5, 62], 2), ([128, 129, 256], 3), ([400, 868, 63, 810], 2),
                                        ([2083], 1), ([300] * 9, 20)])
def test_attn_varlen(svr2lib, lens, heads):
    total = sum(lens)
    q, k, v = (bf(rnd(total, heads, 128, seed=s)) for s in (1, 2, 3))
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    out = svr2lib.attn_varlen(q, k, v, cu, max(lens))
    ref = torch.empty_like(out, dtype=torch.float32)
    o = 0
    for n in lens:
        qi, ki, vi = (x[o:o + n].float().permute(1, 0, 2) for x in (q, k, v))
        ref[o:o + n] = F.scaled_dot_product_attention(qi[None], ki[None], vi[None])[0].permute(1, 0, 2)
        o += n
    assert_close(out, ref, 1e-2, f"attn {lens}")
    # fused scatter (window_reverse)
    perm = torch.randperm(total, device=DEV).int()
    out2 = torch.zeros_like(out)
    svr2lib.attn_varlen(q, k, v, cu, max(lens), out=out2, out_row_map=perm)
    assert torch.equal(out2[perm.long()], out)


# ------------------------------------------------------------------ conv3d
def _to_ndhwc(x_ncdhw, halo):
    x = x_ncdhw[0].permute(1, 2, 3, 0)  # T,H,W,C
    if halo:
        x = torch.cat([x[:1]] * halo + [x], 0)
    return bf(x).contiguous()


@pytest.mark.parametrize("Cin,Cout,k,st,shw,T,H,W", [
    (64, 128, (3, 3, 3), 1, 1, 3, 20, 36),
    (128, 256, (1, 1, 1), 1, 1, 2, 10, 12),
    (128, 128, (1, 3, 3), 1, 2, 3, 16, 24),
    (128, 128, (3, 3, 3), 2, 2, 5, 16, 24),
    (512, 512, (3, 3, 3), 1, 1, 2, 6, 10),
    (128, 8, (3, 3, 3), 1, 1, 2, 12, 20),
    (256, 32, (3, 3, 3), 1, 1, 1, 4, 6),
    (128, 128, (3, 3, 3), 2, 2, 5, 48, 64),      # swap-AB + pair view
    (128, 128, (1, 3, 3), 1, 2, 2, 40, 72),      # swap-AB, spatial-only downsample
    (256, 128, (3, 3, 3), 1, 1, 2, 30, 44),      # swap-AB, ragged tile edges
    (128, 128, (1, 1, 1), 1, 1, 2, 24, 40),      # swap-AB 1x1x1
    (128, 128, (3, 3, 3), 1, 1, 3, 9, 256),      # W-reuse kernel: one 256-pixel row segment per tile, 3 kw taps per load
    (256, 128, (3, 3, 3), 1, 1, 2, 5, 520),      # W-reuse: three segments, ragged last one (8 valid pixels)
    (128, 96, (1, 3, 3), 1, 1, 2, 6, 512),       # W-reuse: kt = 1, Cout < 128
    (128, 128, (1, 3, 3), 1, 2, 2, 8, 1024),     # 256 x 1 tiles on the generic swap-AB kernel (stride 2: not eligible)
    (256, 256, (3, 3, 3), 1, 1, 2, 6, 256),      # CTA-pair W-reuse mainloop: 128-pixel row segments, A descriptors shifted by kw rows
    (256, 512, (3, 3, 3), 1, 1, 1, 3, 128),      # pair W-reuse: two n-tiles, odd number of m-tiles (phantom tile)
    (512, 256, (1, 3, 3), 1, 1, 2, 5, 200),      # pair W-reuse: kt = 1, ragged second segment
    (256, 256, (3, 3, 3), 2, 2, 3, 8, 512),      # 128 x 1 tiles on the generic pair kernel (stride 2: not eligible)
])
def test_conv3d(svr2lib, Cin, Cout, k, st, shw, T, H, W):
    if W >= 128:
        svr2lib.load().svr2_set_conv_wreuse(2)        # W-reuse tiles also where the last row segment is mostly empty
    x = rnd(1, Cin, T, H, W, seed=1)
    w = rnd(Cout, Cin, *k, std=(Cin * k[0] * k[1] * k[2]) ** -0.5, seed=2)
    b = rnd(Cout, seed=3)
    halo = k[0] - 1
    xb, wb, bb = bf(x).float(), bf(w).float(), bf(b).float()
    xp = torch.cat([xb[:, :, :1]] * halo + [xb], 2) if halo else xb
    if shw == 2:
        ref = F.conv3d(F.pad(xp, (0, 1, 0, 1)), wb, bb, stride=(st, 2, 2))
    else:
        ref = F.conv3d(xp, wb, bb, stride=(st, 1, 1), padding=(0, k[1] // 2, k[2] // 2))
    T_out, Ho, Wo = ref.shape[2:]
    x_nd = _to_ndhwc(x, halo)
    w_k = bf(w.permute(0, 2, 3, 4, 1).reshape(Cout, -1)).contiguous()
    res = bf(rnd(T_out, Ho, Wo, Cout, seed=4))
    y = torch.zeros(2 + T_out, Ho, Wo, Cout, device=DEV, dtype=torch.bfloat16)
    svr2lib.conv3d(x_nd, T + halo, H, W, Cin, w_k, Cout, k, st, shw, 1 if (shw == 1 and k[1] == 3) else 0, T_out, y,
                   bias=bf(b), residual=torch.cat([res[:1], res[:1], res], 0).contiguous(), out_t_pad=2,
                   out_dup_head=1)
    ref_nd = bf(bf(ref[0].permute(1, 2, 3, 0)).float() + res.float())
    svr2lib.load().svr2_set_conv_wreuse(1)
    assert_close(y[2:], ref_nd, 4e-3, "conv3d body")
    assert torch.equal(y[0], y[2]) and torch.equal(y[1], y[2]), "halo frames must replicate frame 0"
    if W >= 128 and shw == 1 and k[1] == 3:           # the generic kernel on the same problem: same products,
        y0 = torch.zeros_like(y)                      # another summation order of the fp32 accumulation
        svr2lib.load().svr2_set_conv_wreuse(0)
        try:
            svr2lib.conv3d(x_nd, T + halo, H, W, Cin, w_k, Cout, k, st, shw, 1, T_out, y0, bias=bf(b),
                           residual=torch.cat([res[:1], res[:1], res], 0).contiguous(), out_t_pad=2, out_dup_head=1)
        finally:
            svr2lib.load().svr2_set_conv_wreuse(1)
        assert_close(y, y0, 2e-3, "W-reuse kernel vs generic swap-AB kernel")


@pytest.mark.parametrize("Cin,C2,Cout,T,H,W", [
    (128, 256, 128, 2, 30, 44),     # decoder up3.res0: swap-AB, ragged tile edges
    (256, 512, 256, 3, 24, 40),     # decoder up2.res0: CTA pair
    (256, 128, 256, 2, 17, 33),     # encoder down1.res0 (odd sizes)
    (512, 256, 512, 1, 9, 16),      # encoder down2.res0, single frame
    (128, 256, 128, 2, 7, 512),     # W-reuse kernel with the shortcut's extra k-blocks
    (256, 512, 256, 2, 5, 256),     # CTA-pair W-reuse mainloop with the shortcut's extra k-blocks
])
def test_conv3d_fused_shortcut(svr2lib, Cin, C2, Cout, T, H, W):
    """conv2(h) + conv_shortcut(x) as one contraction over [h ; x] (ResnetBlock3D, attn_video_vae.py:311-362) vs torch:
    fp32 reference of the two convolutions on bf16-rounded operands; statistics slots returned."""
    import ctypes
    h = rnd(1, Cin, T, H, W, seed=1)
    x2 = rnd(1, C2, T, H, W, seed=2)
    w = rnd(Cout, Cin, 3, 3, 3, std=(27 * Cin) ** -0.5, seed=3)
    wsc = rnd(Cout, C2, 1, 1, 1, std=C2 ** -0.5, seed=4)
    b, bsc = rnd(Cout, seed=5), rnd(Cout, seed=6)
    hb, xb = bf(h).float(), bf(x2).float()
    hp = torch.cat([hb[:, :, :1]] * 2 + [hb], 2)
    ref = F.conv3d(hp, bf(w).float(), None, padding=(0, 1, 1)) + F.conv3d(xb, bf(wsc).float(), None)
    bsum = bf(bf(b).float() + bf(bsc).float())
    ref = ref + bsum.float().view(1, -1, 1, 1, 1)
    h_nd, x_nd = _to_ndhwc(h, 2), _to_ndhwc(x2, 0)
    w_cat = torch.cat([bf(w.permute(0, 2, 3, 4, 1).reshape(Cout, -1)), bf(wsc.reshape(Cout, C2))], 1).contiguous()
    y = torch.zeros(2 + T, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    args = (svr2lib.ptr(h_nd), T + 2, H, W, Cin, svr2lib.ptr(w_cat), Cout, 3, 3, 3, T, svr2lib.ptr(bsum),
            svr2lib.ptr(x_nd), C2, svr2lib.ptr(y), 2, 1)
    slots = ctypes.c_int(0)
    assert svr2lib.load().svr2_conv3d_shortcut_stats_bf16(*args, None, 0, ctypes.byref(slots), svr2lib.stream()) == 0
    part = torch.zeros(T * slots.value * (Cout // 8) * 4, device=DEV, dtype=torch.float32)
    svr2lib.call("svr2_conv3d_shortcut_stats_bf16", *args, svr2lib.ptr(part), part.numel() * 4, ctypes.byref(slots),
                 svr2lib.stream())
    assert_close(y[2:], ref[0].permute(1, 2, 3, 0), 4e-3, "conv + fused shortcut")
    assert torch.equal(y[0], y[2]) and torch.equal(y[1], y[2])
    # the statistics the epilogue emitted: per-frame sums of the stored values
    sums = part.view(T, slots.value, Cout // 8, 4)[..., [0, 2]].sum((1, 2, 3))
    assert_close(sums, y[2:].float().sum((1, 2, 3)), 2e-3, "epilogue statistics (sum)")


@pytest.mark.parametrize("C,temporal,F_,H,W", [(256, 0, 3, 6, 10), (512, 1, 3, 4, 6), (512, 1, 1, 4, 6),
                                                # W % 32 == 0: the TMA shuffle-store path (incl. a ragged last m-tile)
                                                (256, 0, 2, 5, 32), (512, 1, 3, 3, 64), (256, 1, 2, 7, 96)])
def test_upsample_shuffle(svr2lib, C, temporal, F_, H, W):
    z = 2 if temporal else 1
    r = 4 * z
    x = rnd(1, C, F_, H, W, seed=1)
    w = rnd(r * C, C, std=C ** -0.5, seed=2)
    b = rnd(r * C, seed=3)
    xb, wb, bb = bf(x).float(), bf(w).float(), bf(b).float()
    y = F.conv3d(xb, wb.view(r * C, C, 1, 1, 1), bb)
    y = y.view(1, 2, 2, z, C, F_, H, W).permute(0, 4, 5, 3, 6, 1, 7, 2).reshape(1, C, F_ * z, 2 * H, 2 * W)
    if temporal:
        y = torch.cat([y[:, :, :1], y[:, :, 2:]], 2)
    T_out = y.shape[2]
    out = torch.zeros(2 + T_out, 2 * H, 2 * W, C, device=DEV, dtype=torch.bfloat16)
    x_nd = _to_ndhwc(x, 0)
    w16, b16 = bf(w).contiguous(), bf(b)          # keep the operands alive across the launch
    svr2lib.call("svr2_upsample_shuffle_bf16", svr2lib.ptr(x_nd), F_, H, W, C, svr2lib.ptr(w16),
                 svr2lib.ptr(b16), temporal, 1, svr2lib.ptr(out), 2, 1, svr2lib.stream())
    assert_close(out[2:], y[0].permute(1, 2, 3, 0), 4e-3, "upsample shuffle")
    assert torch.equal(out[0], out[2]) and torch.equal(out[1], out[2])


# ------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("dim", [256, 2560, 3072])
def test_rmsnorm_ada(svr2lib, dim):
    x = bf(rnd(333, dim, seed=1))
    scale, shift, wt = rnd(dim, seed=2) * 0.1 + 1, rnd(dim, seed=3) * 0.1, rnd(dim, seed=4) * 0.1 + 1
    r = x.float() / torch.sqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-5)
    out = svr2lib.rmsnorm_ada(x, scale, shift, mode=0)
    assert_close(out, r * scale + shift, 3e-3, "mode0")
    out = svr2lib.rmsnorm_ada(x, scale, shift, weight=wt, mode=0)
    assert_close(out, r * wt * scale + shift, 3e-3, "mode0+weight")
    out = svr2lib.rmsnorm_ada(x, scale, shift, mode=1)
    ref = bf(bf(bf(r).float() * scale).float() + shift)
    assert_close(out, ref, 3e-3, "mode1")


def test_qk_norm_rope_window(svr2lib):
    heads, L, l, nf = 3, 50, 7, 21
    inner = heads * 128
    qkv_v, qkv_t = bf(rnd(L, 3 * inner, seed=1)), bf(rnd(l, 3 * inner, seed=2))
    wq_v, wk_v, wq_t, wk_t = (rnd(128, seed=s) * 0.1 + 1 for s in (3, 4, 5, 6))
    R = 40
    ang = rnd(R, nf, seed=7)
    cos_t, sin_t = ang.cos().contiguous(), ang.sin().contiguous()
    g = torch.Generator().manual_seed(0)
    total = 90
    src = torch.randint(0, L, (total,), generator=g)
    is_txt = torch.rand(total, generator=g) < 0.3
    src = torch.where(is_txt, -(torch.randint(0, l, (total,), generator=g) + 1), src).int().to(DEV)
    rope = torch.randint(0, R, (total, 3), generator=g).int()
    rope[::5] = -1
    rope = rope.to(DEV)
    q = torch.empty(total, heads, 128, device=DEV, dtype=torch.bfloat16)
    k, v = torch.empty_like(q), torch.empty_like(q)
    svr2lib.call("svr2_qk_norm_rope_window_bf16", *(svr2lib.ptr(t) for t in (qkv_v, qkv_t, src, rope.contiguous(),
                 cos_t, sin_t)), nf, *(svr2lib.ptr(t) for t in (wq_v, wk_v, wq_t, wk_t)), 1e-5, total, heads,
                 svr2lib.ptr(q), svr2lib.ptr(k), svr2lib.ptr(v), svr2lib.stream())
    srcl = src.long()
    rows = torch.where((srcl < 0)[:, None], qkv_t[(-srcl - 1).clamp_min(0)].float(), qkv_v[srcl.clamp_min(0)].float())
    rows = rows.view(total, 3, heads, 128)
    for which, (wv, wt), got in ((0, (wq_v, wq_t), q), (1, (wk_v, wk_t), k)):
        x = rows[:, which]
        x = x / torch.sqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)
        x = x * torch.where((srcl < 0)[:, None, None], wt, wv)
        c = torch.ones(total, 128, device=DEV)
        s = torch.zeros(total, 128, device=DEV)
        for ax in range(3):
            idx = rope[:, ax].long()
            cc = torch.where((idx >= 0)[:, None], cos_t[idx.clamp_min(0)], torch.ones(1, device=DEV))
            ss = torch.where((idx >= 0)[:, None], sin_t[idx.clamp_min(0)], torch.zeros(1, device=DEV))
            c[:, ax * 2 * nf:(ax + 1) * 2 * nf] = cc.repeat_interleave(2, -1)
            s[:, ax * 2 * nf:(ax + 1) * 2 * nf] = ss.repeat_interleave(2, -1)
        x1, x2 = x[..., 0::2], x[..., 1::2]
        rot = torch.stack((-x2, x1), -1).reshape(x.shape)
        ref = x * c[:, None] + rot * s[:, None]
        assert_close(got, ref, 3e-3, "qk" + str(which))
    assert torch.equal(v.float(), rows[:, 2])


@pytest.mark.parametrize("variant,heads,geom", [("3b", 2, (3, 20, 36)), ("3b", 4, (5, 34, 60)), ("7b", 2, (2, 20, 36)),
                                                ("3b", 20, (1, 10, 14))])
def test_linear_qkv_rope_fused_vs_unfused(pkg, svr2lib, variant, heads, geom):
    """QKV GEMM with q/k RMSNorm + RoPE + window scatter in its epilogue (svr2_linear_qkv_rope_bf16) + the row-subset
    kernel for the text rows, against the two-kernel path (svr2_linear_bf16 -> svr2_qk_norm_rope_window_bf16) on the real
    window layouts (regular and shifted): v bit-equal, q/k equal up to the summation order of the per-head RMS."""
    import importlib
    dit = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.dit")
    T, Hp, Wp = geom
    l, d, inner = 58, heads * 128, heads * 128
    L = T * Hp * Wp
    a_v, a_t = bf(rnd(L, d, seed=1)), bf(rnd(l, d, seed=2))
    w_v, w_t = bf(rnd(3 * inner, d, std=d ** -0.5, seed=3)), bf(rnd(3 * inner, d, std=d ** -0.5, seed=4))
    nq_v, nk_v, nq_t, nk_t = (rnd(128, seed=s) * 0.1 + 1 for s in (5, 6, 7, 8))
    nqk = torch.cat([nq_v, nk_v]).contiguous()
    if variant == "3b":
        freqs = (1.0 / (10000 ** (torch.arange(0, 42, 2)[:21].float() / 42))).half()
    else:
        freqs = (torch.linspace(1.0, 128.0, 10) * math.pi).half()
    for shifted in (False, True):
        lay, size_rows = dit.build_layout(T, Hp, Wp, l, shifted, variant, DEV)
        c, s = dit.rope_tables(freqs, variant, int(lay.row_rope.max().item()) + 1, size_rows)
        cos_t, sin_t = c.to(DEV), s.to(DEV)
        nf = cos_t.shape[1]
        P = svr2lib.ptr
        qkv_v, qkv_t = svr2lib.linear(a_v, w_v), svr2lib.linear(a_t, w_t)
        ref = [torch.zeros(lay.total, heads, 128, device=DEV, dtype=torch.bfloat16) for _ in range(3)]
        svr2lib.call("svr2_qk_norm_rope_window_bf16", P(qkv_v), P(qkv_t), P(lay.row_src), P(lay.row_rope), P(cos_t),
                     P(sin_t), nf, P(nq_v), P(nk_v), P(nq_t), P(nk_t), 1e-5, lay.total, heads, *(P(t) for t in ref),
                     svr2lib.stream())
        got = [torch.zeros_like(t) for t in ref]
        svr2lib.call("svr2_linear_qkv_rope_bf16", P(a_v), d, P(w_v), d, L, heads, d, P(lay.tok_dst), P(lay.tok_rope),
                     P(cos_t), P(sin_t), nf, P(nqk), 1e-5, *(P(t) for t in got), svr2lib.stream())
        svr2lib.call("svr2_qk_norm_rope_rows_bf16", None, P(qkv_t), P(lay.row_src), P(lay.row_rope), P(cos_t), P(sin_t),
                     nf, P(nq_v), P(nk_v), P(nq_t), P(nk_t), 1e-5, P(lay.txt_rows), lay.txt_rows.numel(), heads,
                     *(P(t) for t in got), svr2lib.stream())
        assert torch.equal(got[2], ref[2]), "v rows must be bit-equal"
        for name, g_, r_ in (("q", got[0], ref[0]), ("k", got[1], ref[1])):
            dlt = (g_.float() - r_.float()).abs()
            assert (dlt == 0).float().mean() > 0.98 and dlt.max() <= 2 ** -6 * r_.abs().max().item(), \
                f"{name} shifted={shifted}: {(dlt == 0).float().mean():.4f} equal, max {dlt.max():.4f}"


@pytest.mark.parametrize("C,hw,frames,silu", [(128, 24 * 36, 3, 1), (256, 1000, 2, 1), (512, 77, 2, 0),
                                              (128, 300 * 200, 2, 1), (512, 20000, 1, 1)])
def test_groupnorm(svr2lib, C, hw, frames, silu):
    x = bf(rnd(frames, hw, C, seed=1) * 2 + 0.5)
    gamma, beta = bf(rnd(C, seed=2) * 0.1 + 1), bf(rnd(C, seed=3) * 0.1)
    y = torch.zeros(2 + frames, hw, C, device=DEV, dtype=torch.bfloat16)
    need = svr2lib.load().svr2_groupnorm_scratch_bytes(frames, hw, C)
    stats = torch.zeros(need // 8 + 1, device=DEV, dtype=torch.float64)
    args = (svr2lib.ptr(x), svr2lib.ptr(y), frames, hw, C, svr2lib.ptr(gamma), svr2lib.ptr(beta), 1e-6, silu, 2, 1,
            svr2lib.ptr(stats), stats.numel() * 8, svr2lib.stream())
    svr2lib.call("svr2_groupnorm_bf16", *args)
    y_first = y.clone()
    svr2lib.call("svr2_groupnorm_bf16", *args)
    assert torch.equal(y, y_first), "groupnorm must be bit-reproducible"
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, gamma.float(), beta.float(), 1e-6)
    ref = bf(ref).float()
    if silu:
        ref = F.silu(ref)
    assert_close(y[2:], ref.permute(0, 2, 1), 3e-3, "groupnorm")
    assert torch.equal(y[0], y[2]) and torch.equal(y[1], y[2])


def test_softmax_transpose_misc(svr2lib):
    s = rnd(37, 1000, seed=1) * 3
    p = torch.empty(37, 1000, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_softmax_rows_bf16", svr2lib.ptr(s), 1000, svr2lib.ptr(p), 1000, 37, 1000, svr2lib.stream())
    assert_close(p, torch.softmax(s, -1), 3e-3, "softmax")
    a = bf(rnd(70, 130, seed=2))
    t = torch.empty(130, 72, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_transpose_bf16", svr2lib.ptr(a), 130, svr2lib.ptr(t), 72, 70, 130, svr2lib.stream())
    assert torch.equal(t[:, :70], a.T)
    # txt mean
    x = bf(rnd(9, 58, 256, seed=3))
    o = torch.empty(58, 256, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_txt_window_mean_bf16", svr2lib.ptr(x), svr2lib.ptr(o), 9, 58, 256, svr2lib.stream())
    assert_close(o, x.float().mean(0), 3e-3, "txt mean")
    # patchify / unpatchify
    T, H, W, C = 2, 6, 8, 33
    vid = bf(rnd(T * H * W, C, seed=4))
    pt = torch.empty(T * 3 * 4, 192, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_patchify_bf16", svr2lib.ptr(vid), svr2lib.ptr(pt), T, H, W, C, 192, svr2lib.stream())
    ref = vid.view(T, 3, 2, 4, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(T * 12, 4 * C)
    assert torch.equal(pt[:, :132], ref) and pt[:, 132:].abs().max() == 0
    back = torch.empty(T * H * W, 16, device=DEV, dtype=torch.bfloat16)
    v64 = bf(rnd(T * 12, 64, seed=5))
    svr2lib.call("svr2_unpatchify_bf16", svr2lib.ptr(v64), 64, svr2lib.ptr(back), T, H, W, 16, svr2lib.stream())
    ref = v64.view(T, 3, 4, 2, 2, 16).permute(0, 1, 3, 2, 4, 5).reshape(T * H * W, 16)
    assert torch.equal(back, ref)


def test_layout_and_im2col(svr2lib):
    C, T, H, W = 3, 2, 6, 8
    x = rnd(C, T, H, W, seed=1)
    out = torch.full((2 + T, H, W, 8), 7.0, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_ncdhw_to_ndhwc_bf16", svr2lib.ptr(x), 0, C, T, H, W, svr2lib.ptr(out), 8, 2, 1.0,
                 svr2lib.stream())
    assert torch.equal(out[2:, ..., :3], bf(x).permute(1, 2, 3, 0)) and out[..., 3:].abs().max() == 0
    assert torch.equal(out[0], out[2]) and torch.equal(out[1], out[2])
    back = torch.empty(C, T, H, W, device=DEV, dtype=torch.float32)
    svr2lib.call("svr2_ndhwc_to_ncdhw", svr2lib.ptr(out[2:]), 8, C, T, H, W, svr2lib.ptr(back), 0, svr2lib.stream())
    assert torch.equal(back, bf(x).float())
    col = torch.empty(T * H * W, 128, device=DEV, dtype=torch.bfloat16)
    svr2lib.call("svr2_im2col3_bf16", svr2lib.ptr(out), T, H, W, C, 8, svr2lib.ptr(col), 128, svr2lib.stream())
    xp = F.pad(out[..., :3].float().permute(3, 0, 1, 2)[None], (1, 1, 1, 1))  # halo already in T
    ref = xp.unfold(2, 3, 1).unfold(3, 3, 1).unfold(4, 3, 1)  # 1,C,T,H,W,kt,kh,kw
    ref = ref[0].permute(1, 2, 3, 4, 5, 6, 0).reshape(T * H * W, 81)
    assert torch.equal(col[:, :81].float(), ref) and col[:, 81:].abs().max() == 0


@pytest.mark.parametrize("M,n", [(300, 160), (1000, 2052), (129, 36)])
def test_two_pass_attention_probabilities(svr2lib, M, n):
    """EPI_ROWSTAT + rowstat_combine + EPI_PEXP == softmax(q k^T * scale) (VAE mid-block attention)."""
    d = 512
    q, k = bf(rnd(M, d, seed=1)), bf(rnd(n, d, seed=2))
    scale = d ** -0.5
    s2 = scale * 1.4426950408889634
    slots = svr2lib.load().svr2_rowstat_slots(n)
    part = torch.empty(M, 2 * slots, device=DEV, dtype=torch.float32)
    svr2lib.linear(q, k, epi=svr2lib.EPI_ROWSTAT, out=part, out_scale=s2)
    lse = torch.empty(M, device=DEV, dtype=torch.float32)
    svr2lib.call("svr2_rowstat_combine", svr2lib.ptr(part), slots, slots, svr2lib.ptr(lse), M, svr2lib.stream())
    ldn = (n + 7) // 8 * 8
    P = torch.zeros(M, ldn, device=DEV, dtype=torch.bfloat16)
    k_pad = torch.zeros(ldn, d, device=DEV, dtype=torch.bfloat16)   # bf16 output needs N % 8 == 0
    k_pad[:n] = k
    svr2lib.linear(q, k_pad, epi=svr2lib.EPI_PEXP, gate=lse, out=P, out_scale=s2)
    S = (q.float() @ k.float().T) * scale
    assert_close(lse, torch.logsumexp(S, -1) * 1.4426950408889634, 1e-4, "lse2")
    assert_close(P[:, :n], torch.softmax(S, -1), 6e-3, "probabilities")
    assert (P[:, :n].float().sum(-1) - 1).abs().max() < 2e-2


@pytest.mark.parametrize("Cin,Cout,T,H,W", [(64, 128, 2, 20, 36), (128, 256, 2, 19, 30), (256, 512, 1, 12, 20)])
def test_conv_epilogue_groupnorm_stats(svr2lib, Cin, Cout, T, H, W):
    """svr2_conv3d_stats_bf16 + svr2_groupnorm_from_stats_bf16 == conv followed by per-frame GroupNorm + SiLU."""
    import ctypes
    x = rnd(1, Cin, T, H, W, seed=1)
    w = rnd(Cout, Cin, 3, 3, 3, std=(Cin * 27) ** -0.5, seed=2)
    b = rnd(Cout, seed=3)
    x_nd = _to_ndhwc(x, 2)
    w_k = bf(w.permute(0, 2, 3, 4, 1).reshape(Cout, -1)).contiguous()
    y = torch.zeros(T, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    b16 = bf(b)                                   # must outlive the launch (a temporary's block gets reused by `part`)
    args = (svr2lib.ptr(x_nd), T + 2, H, W, Cin, svr2lib.ptr(w_k), Cout, 3, 3, 3, 1, 1, 1, T, svr2lib.EPI_BIAS,
            svr2lib.ptr(b16), None, svr2lib.ptr(y), 0, 0, Cout)
    slots = ctypes.c_int(0)
    assert svr2lib.load().svr2_conv3d_stats_bf16(*args, None, 0, ctypes.byref(slots), svr2lib.stream()) == 0
    part = torch.full((T * slots.value * (Cout // 8) * 4,), float("nan"), device=DEV)
    svr2lib.call("svr2_conv3d_stats_bf16", *args, svr2lib.ptr(part), part.numel() * 4, ctypes.byref(slots),
                 svr2lib.stream())
    assert torch.isfinite(part).all(), "every partial slot must be written"
    gamma, beta = bf(rnd(Cout, seed=4) * 0.1 + 1), bf(rnd(Cout, seed=5) * 0.1)
    z = torch.zeros(2 + T, H * W, Cout, device=DEV, dtype=torch.bfloat16)
    coef = torch.empty(T * Cout * 2, device=DEV)
    svr2lib.call("svr2_groupnorm_from_stats_bf16", svr2lib.ptr(y), svr2lib.ptr(z), T, H * W, Cout, svr2lib.ptr(gamma),
                 svr2lib.ptr(beta), 1e-6, 1, 2, 1, svr2lib.ptr(part), slots.value, svr2lib.ptr(coef), svr2lib.stream())
    ref = F.group_norm(y.float().view(T, H * W, Cout).permute(0, 2, 1), 32, gamma.float(), beta.float(), 1e-6)
    ref = F.silu(bf(ref).float()).permute(0, 2, 1)
    assert_close(z[2:], ref, 3e-3, "fused-stats groupnorm")
    assert torch.equal(z[0], z[2]) and torch.equal(z[1], z[2])
