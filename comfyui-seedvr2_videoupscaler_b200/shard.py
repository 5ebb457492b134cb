"""Clip-parallel sharding across the GPUs of one box.

The path shards by clip: every rank upscales a contiguous frame range with its own
full copy of the weights (no data-path collective), then ONE all-gather returns the
decoded frames — replacing the reference's mp.Queue + shared-memory + numpy hand-off
(``inference_cli.py:1100, 1227-1232``).  The partition is the reference's:
``total // n`` frames per rank, +1 for the first ``total % n`` ranks, plus
``temporal_overlap`` extra frames on all but the last rank
(``inference_cli.py:1166-1176``; ``partition_preloaded`` is the variant for frames already in memory, ``:1196-1213``).
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def partition_frames(total: int, n: int, overlap: int = 0, start: int = 0) -> List[Tuple[int, int]]:
    """[start, end) per rank, reference order (inference_cli.py:1166-1193)."""
    base, rem = total // n, total % n
    out, cur = [], start
    for idx in range(n):
        cnt = base + (1 if idx < rem else 0)
        end = cur + cnt
        if idx < n - 1 and overlap > 0:
            end = min(end + overlap, start + total)
        out.append((cur, end))
        cur += cnt
    return out


def partition_preloaded(total: int, n: int, overlap: int = 0, batch_size: int = 1) -> List[Tuple[int, int]]:
    """[start, end) per rank when the whole frame tensor is already in memory (inference_cli.py:1196-1213): without
    overlap ``torch.chunk`` (ceil(total / n) frames per rank, the tail ranks may get fewer or none); with overlap,
    chunks of ``total // n + overlap`` frames rounded up to a multiple of ``batch_size``, stepping by that minus the
    overlap, the last rank running to the end."""
    if overlap > 0 and n > 1:
        cwo = total // n + overlap
        if batch_size > 1:
            cwo = (cwo + batch_size - 1) // batch_size * batch_size
        base = cwo - overlap
        out = []
        for i in range(n):
            s = i * base
            e = total if i == n - 1 else min(s + cwo, total)
            out.append((min(s, total), max(min(s, total), e)))
        return out
    size = -(-total // n)          # torch.chunk
    out, cur = [], 0
    while cur < total:
        out.append((cur, min(cur + size, total)))
        cur += size
    return out


def gather_frames(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """All-gather decoded frames (T_r, H, W, C) of every rank into rank order.
    Ranks may hold different frame counts: shards are padded to max(counts) for one
    equal-sized NCCL/gloo all_gather and trimmed afterwards."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    tmax = max(counts)
    T, H, W, C = local.shape
    buf = local
    if T < tmax:
        buf = torch.cat([local, local.new_zeros(tmax - T, H, W, C)], 0)
    out = torch.empty(world, tmax, H, W, C, device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out.view(-1), buf.reshape(-1).contiguous(), group=group)
    return torch.cat([out[r, :counts[r]] for r in range(world)], 0)


def blend_weights(overlap: int, dtype=torch.bfloat16):
    """Per-frame (w_prev, w_cur) of blend_overlapping_frames (generation_utils.py:299-310), computed with the same
    torch ops in the frames' dtype so that every rounding point matches: Hann cross-fade over the middle third for
    overlap >= 3, linear below."""
    if overlap >= 3:
        t = torch.linspace(0.0, 1.0, steps=overlap, dtype=dtype)
        blend_start, blend_end = 1.0 / 3.0, 2.0 / 3.0
        u = ((t - blend_start) / (blend_end - blend_start)).clamp(0.0, 1.0)
        w_prev = 0.5 + 0.5 * torch.cos(torch.pi * u)
    else:
        w_prev = torch.linspace(1.0, 0.0, steps=overlap, dtype=dtype)
    return w_prev, 1.0 - w_prev


def blend_overlap(prev_tail: torch.Tensor, cur_head: torch.Tensor) -> torch.Tensor:
    """Cross-fade of the ``overlap`` frames two neighbouring ranges share (``blend_overlapping_frames``,
    generation_utils.py:284-312): [overlap, H, W, C] each, bf16 (inside the pipeline) or fp32 (the multi-GPU merge),
    on the GPU (one libsvr2 kernel); weights and rounding points follow the frames' dtype."""
    from . import lib
    assert prev_tail.shape == cur_head.shape and prev_tail.is_cuda
    n = prev_tail.shape[0]
    dt = torch.float32 if prev_tail.dtype == torch.float32 else torch.bfloat16
    a, b = prev_tail.to(dt).contiguous(), cur_head.to(dt).contiguous()
    w_prev, w_cur = blend_weights(n, dt)
    wp, wc = w_prev.float().to(a.device), w_cur.float().to(a.device)
    elems = a[0].numel()
    shape = a.shape
    pad = (-elems) % (4 if dt == torch.float32 else 8)     # the kernels move 16 bytes per thread
    if pad:      # odd frame sizes (e.g. a max_resolution cap that rounds to odd H and W): pad each frame's tail
        a = torch.nn.functional.pad(a.reshape(n, elems), (0, pad))
        b = torch.nn.functional.pad(b.reshape(n, elems), (0, pad))
    out = torch.empty_like(a)
    name = "svr2_blend_overlap_f32" if dt == torch.float32 else "svr2_blend_overlap_bf16"
    lib.call(name, lib.ptr(a), lib.ptr(b), lib.ptr(out), lib.ptr(wp), lib.ptr(wc), n, elems + pad, lib.stream(),
             nbytes=3.0 * a.numel() * a.element_size())
    return out[:, :elems].reshape(shape) if pad else out


def merge_shards(chunks: List[torch.Tensor], overlap: int, blend=None) -> torch.Tensor:
    """Concatenate the per-rank results in rank order, cross-fading the ``overlap`` frames a chunk shares with the
    accumulated result (inference_cli.py:1241-1274; fp32 like the reference).  ``blend(prev_tail, cur_head)``
    defaults to the libsvr2 kernel."""
    blend = blend or blend_overlap
    chunks = [c.float() for c in chunks]
    if overlap <= 0 or len(chunks) == 1:
        return torch.cat(chunks, 0)
    result = chunks[0]
    for c in chunks[1:]:
        if c.shape[0] > overlap and result.shape[0] >= overlap:
            blended = blend(result[-overlap:], c[:overlap])
            result = torch.cat([result[:-overlap], blended, c[overlap:]], 0)
        elif c.shape[0] > overlap:       # chunk too small to blend into: append its non-overlapping part
            result = torch.cat([result, c[overlap:]], 0)
    return result
