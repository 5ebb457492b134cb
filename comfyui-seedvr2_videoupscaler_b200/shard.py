"""Clip-parallel sharding across the GPUs of one box.

The path shards by clip: every rank upscales a contiguous frame range with its own
full copy of the weights (no data-path collective), then ONE all-gather returns the
decoded frames — replacing the reference's mp.Queue + shared-memory + numpy hand-off
(``inference_cli.py:1100, 1227-1232``).  The partition is the reference's:
``total // n`` frames per rank, +1 for the first ``total % n`` ranks, plus
``temporal_overlap`` extra frames on all but the last rank
(``inference_cli.py:1166-1176``).
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


def blend_overlap(prev_tail: torch.Tensor, cur_head: torch.Tensor) -> torch.Tensor:
    """Linear cross-fade of the overlapping frames of two neighbouring ranges
    (generation_utils.py:284-312 uses a Hann/linear window; linear here)."""
    n = prev_tail.shape[0]
    w = torch.linspace(0, 1, n + 2, device=prev_tail.device, dtype=torch.float32)[1:-1].view(n, 1, 1, 1)
    return (prev_tail.float() * (1 - w) + cur_head.float() * w).to(prev_tail.dtype)
