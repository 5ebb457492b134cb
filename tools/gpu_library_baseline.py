"""The bar north_star names: the reference's own GPU libraries on the same box, timed beside the engine.

The reference is pure Python over PyTorch library kernels: cuBLASLt (`nn.Linear`, mmattn.py:56-59), cuDNN Conv3d
(`InflatedCausalConv3d`, causal_inflation_lib.py:101-104,143), torch SDPA per window (attention.py:27-64) or flash-attn-2
varlen (compatibility.py:287-330).  The oracle's "ref_bf16" flow (oracle/*.py, pinned to the reference modules) executes
exactly those library calls on CUDA tensors, so timing it per phase on the GPU IS the reference's GPU path with weights
resident and no host bounce — the "reference-vs-engine" bar of BASELINE.md §3.  Op level: the three library kernels
SURVEY §2b marks as the ones to beat, on the shapes the 4K shard runs.

    python tools/gpu_library_baseline.py [--ops] [--phases cfg2|4k_shard] [--out gpurun_out/gpu_library_baseline.json]

`bench.py` imports `op_level()` / `phase_level()` and puts the result into its JSON line as `gpu_library_baseline`.
Test infrastructure: nothing here is on the product path.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(fn, warmup=2, iters=5):
    """median CUDA-event time (ms) of fn() on the current stream"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def _rnd(*s, scale=1.0, dev="cuda"):
    return (torch.randn(*s, device=dev, dtype=torch.float32) * scale).to(torch.bfloat16)


def op_level(dev="cuda"):
    """engine C-ABI op vs the library kernel the reference calls, same shape, same dtype, CUDA events."""
    from svr2_import import load_package
    load_package()
    lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
    out = {}

    # ---- windowed attention, 4K shard: 243 windows x (405 + 58) tokens x 20 heads (dit_3b/attention.py:114-148)
    for name, lens in (("attn_243x463x20", [463] * 243), ("attn_400x2083x20", [2083] * 400)):
        total = sum(lens)
        q, k, v = (_rnd(total, 20, 128) for _ in range(3))
        cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
        o = torch.empty_like(q)
        flops = 4.0 * sum(n * n for n in lens) * 128 * 20
        row = {"flops": flops}
        row["engine_ms"] = _time(lambda: lib.attn_varlen(q, k, v, cu, max(lens), out=o))
        try:
            from flash_attn import flash_attn_varlen_func
            row["flash_attn2_ms"] = _time(lambda: flash_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens)))
        except Exception as ex:   # noqa: BLE001
            row["flash_attn2_ms"] = None
            row["flash_attn2_error"] = f"{type(ex).__name__}: {ex}"[:200]
        # the reference's default: a Python loop of SDPA calls, one per window (attention.py:39-62), incl. its cu_seqlens.cpu()
        n_loop = min(len(lens), 243)

        def sdpa_loop():
            cpu = cu.cpu()
            outs = []
            for i in range(n_loop):
                a, b = int(cpu[i]), int(cpu[i + 1])
                qi, ki, vi = (t[a:b].permute(1, 0, 2).unsqueeze(0) for t in (q, k, v))
                outs.append(F.scaled_dot_product_attention(qi, ki, vi).squeeze(0).permute(1, 0, 2))
            return torch.cat(outs, 0)
        ms = _time(sdpa_loop, warmup=1, iters=3)
        row["sdpa_loop_ms"] = ms * len(lens) / n_loop
        for key in ("engine_ms", "flash_attn2_ms", "sdpa_loop_ms"):
            if row.get(key):
                row[key.replace("_ms", "_tflops")] = flops / row[key] / 1e9
        out[name] = row
        del q, k, v, o

    # ---- Linear, 4K shard SwiGLU in-projection: 97200 x 13824 x 2560 (mlp.py:56-61) and the attention out-projection
    for name, (M, N, K) in (("linear_97200x13824x2560", (97200, 13824, 2560)), ("linear_97200x2560x2560", (97200, 2560, 2560)),
                            ("linear_97200x2560x6912", (97200, 2560, 6912))):
        a, w = _rnd(M, K), _rnd(N, K, scale=K ** -0.5)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * M * N * K
        row = {"flops": flops}
        row["engine_ms"] = _time(lambda: lib.linear(a, w, out=y))
        row["cublas_ms"] = _time(lambda: torch.matmul(a, w.t(), out=y))
        row["engine_tflops"], row["cublas_tflops"] = flops / row["engine_ms"] / 1e9, flops / row["cublas_ms"] / 1e9
        out[name] = row
        del a, w, y

    # ---- causal Conv3d (causal_inflation_lib.py:228-248): the two ncu'd shapes.  cuDNN gets its preferred layout
    #      (channels_last_3d) and a pre-padded input, i.e. its time excludes the reference's torch.cat / F.pad copies.
    for name, (C, Co, T, H, W) in (("conv3d_256to256_2x1080x1920", (256, 256, 2, 1080, 1920)),
                                   ("conv3d_128to128_2x2160x3840", (128, 128, 2, 2160, 3840))):
        x = _rnd(T + 2, H, W, C)
        wk = _rnd(Co, 27 * C, scale=(27 * C) ** -0.5)
        b = _rnd(Co)
        y = torch.empty(T, H, W, Co, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * T * H * W * Co * 27 * C
        row = {"flops": flops}
        row["engine_ms"] = _time(lambda: lib.conv3d(x, T + 2, H, W, C, wk, Co, (3, 3, 3), 1, 1, 1, T, y, bias=b))
        xt = x.permute(3, 0, 1, 2)[None]                                    # NCDHW view with channels-last strides
        wt = wk.view(Co, 3, 3, 3, C).permute(0, 4, 1, 2, 3).contiguous(memory_format=torch.channels_last_3d)
        try:
            row["cudnn_channels_last_ms"] = _time(lambda: F.conv3d(xt, wt, b, padding=(0, 1, 1)), warmup=2, iters=3)
        except Exception as ex:   # noqa: BLE001
            row["cudnn_channels_last_ms"] = None
            row["cudnn_channels_last_error"] = f"{type(ex).__name__}: {ex}"[:200]
        try:
            xc, wc = xt.contiguous(), wt.contiguous()                        # the reference's actual layout: NCDHW
            row["cudnn_ncdhw_ms"] = _time(lambda: F.conv3d(xc, wc, b, padding=(0, 1, 1)), warmup=2, iters=3)
            del xc, wc
        except Exception as ex:   # noqa: BLE001
            row["cudnn_ncdhw_ms"] = None
            row["cudnn_ncdhw_error"] = f"{type(ex).__name__}: {ex}"[:200]
        for key in ("engine_ms", "cudnn_channels_last_ms", "cudnn_ncdhw_ms"):
            if row.get(key):
                row[key.replace("_ms", "_tflops")] = flops / row[key] / 1e9
        out[name] = row
        del x, wk, b, y, xt, wt
        torch.cuda.empty_cache()
    return out


def phase_level(workload="cfg2", engine=None, dev="cuda", variant="3b"):
    """encode / DiT / decode of one clip: the engine vs the oracle's ref_bf16 flow (= the reference's library calls:
    cuDNN Conv3d, cuBLASLt Linear under bf16, SDPA per window) on the same GPU, CUDA events, weights resident."""
    from svr2_import import load_package
    pkg = load_package()
    pipeline = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.pipeline")
    from oracle import dit_oracle, vae_oracle
    frames_pad, H, W = {"cfg2": (17, 1088, 1920), "4k_shard": (9, 2160, 3840), "720p": (9, 720, 1280)}[workload]
    if engine is None:
        engine = pipeline.build_synthetic_engine(variant, device=dev)
    cfg = dit_oracle.dit_config(variant)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(3, frames_pad, H, W, generator=g) * 2 - 1).to(dev, torch.bfloat16)
    res = {"workload": f"{frames_pad} frames {H}x{W}, SeedVR2-{variant}", "library_flow":
           "oracle ref_bf16 on CUDA: F.conv3d (cuDNN, bf16 NCDHW) + F.linear (cuBLASLt, weights cast per call like "
           "autocast) + F.scaled_dot_product_attention per window"}

    def ev(fn, iters=2):
        fn()
        torch.cuda.synchronize()
        return _time(fn, warmup=0, iters=iters)

    # ---- engine phases
    lat = engine.vae_encode(x)
    noise = torch.randn(lat.shape, generator=torch.Generator(device=dev).manual_seed(1), device=dev, dtype=torch.bfloat16)
    x0 = engine.inference(noise, lat)
    res["engine_ms"] = {"encode": ev(lambda: engine.vae_encode(x)), "dit": ev(lambda: engine.inference(noise, lat)),
                        "decode": ev(lambda: engine.vae_decode(x0))}
    sys.modules["comfyui_seedvr2_videoupscaler_b200.lib"].release_workspace()   # the library flow needs the engine's resident block
    torch.cuda.empty_cache()
    # ---- library flow (same synthetic weights, regenerated from the seeds build_synthetic_engine uses)
    vae_sd = pkg.weights.synth_vae_state_dict(seed=1235, dtype=torch.float16, device=dev)
    lib_ms = {}
    try:
        lib_ms["encode"] = ev(lambda: vae_oracle.runner_encode(vae_sd, x[None], "ref_bf16"), iters=1)
        torch.cuda.empty_cache()
        lib_ms["decode"] = ev(lambda: vae_oracle.runner_decode(vae_sd, x0, "ref_bf16"), iters=1)
    except torch.OutOfMemoryError as ex:
        lib_ms["vae_error"] = f"OutOfMemoryError: {str(ex)[:120]}"
    del vae_sd
    torch.cuda.empty_cache()
    dit_sd = pkg.weights.synth_dit_state_dict(cfg, seed=1234, dtype=torch.float16, device=dev)
    Tl, h, w_, c = lat.shape
    vid = torch.cat([noise, lat, torch.ones(Tl, h, w_, 1, device=dev, dtype=torch.bfloat16)], -1).view(Tl * h * w_, 2 * c + 1)
    for impl in ("sdpa", "flash_attn"):
        try:
            lib_ms[f"dit_{impl}"] = ev(lambda: dit_oracle.dit_forward(dit_sd, cfg, vid, engine.txt, Tl, h, w_, mode="ref_bf16",
                                                                       attn_impl=impl), iters=1)
        except Exception as ex:   # noqa: BLE001
            lib_ms[f"dit_{impl}_error"] = f"{type(ex).__name__}: {str(ex)[:120]}"
    del dit_sd
    torch.cuda.empty_cache()
    res["library_ms"] = lib_ms
    e = res["engine_ms"]
    if "encode" in lib_ms and "decode" in lib_ms and "dit_sdpa" in lib_ms:
        best_dit = min(v for k, v in lib_ms.items() if k.startswith("dit_") and isinstance(v, float))
        res["speedup"] = {"encode": lib_ms["encode"] / e["encode"], "decode": lib_ms["decode"] / e["decode"],
                          "dit": best_dit / e["dit"],
                          "clip": (lib_ms["encode"] + lib_ms["decode"] + best_dit) / (e["encode"] + e["decode"] + e["dit"])}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", action="store_true")
    ap.add_argument("--phases", default="", help="comma list of cfg2,4k_shard,720p")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_library_baseline.json"))
    args = ap.parse_args()
    res = {}
    if args.ops:
        res["ops"] = op_level()
    eng = None
    for wl in [s for s in args.phases.split(",") if s]:
        if eng is None:
            from svr2_import import load_package
            load_package()
            pipeline = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.pipeline")
            eng = pipeline.build_synthetic_engine("3b")
        res[f"phases_{wl}"] = phase_level(wl, engine=eng)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
