"""CPU, gloo, world_size 2: the N>1 path — reference frame partition + one all-gather of the
decoded frames (ragged shards) reproduces the single-process order."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    from svr2_import import load_package
    load_package()
    import importlib
    shard = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.shard")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard.partition_frames(total, world)
    s, e = parts[rank]
    frames = torch.arange(total, dtype=torch.float32).view(total, 1, 1, 1).expand(total, 2, 3, 3).contiguous()
    local = frames[s:e] * 2 + 1          # stand-in for "upscale my clip"
    out = shard.gather_frames(local, [b - a for a, b in parts])
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world = 7, 2
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = torch.arange(total, dtype=torch.float32).view(total, 1, 1, 1).expand(total, 2, 3, 3) * 2 + 1
    assert torch.equal(out, ref)


def _worker_overlap(rank, world, port, total, overlap, q):
    sys.path.insert(0, ROOT)
    from svr2_import import load_package
    load_package()
    import importlib
    shard = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.shard")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard.partition_frames(total, world, overlap)
    s, e = parts[rank]
    g = torch.Generator().manual_seed(100 + rank)
    local = torch.rand(e - s, 4, 6, 3, generator=g)          # every rank "upscales" its (overlapping) range
    counts = [b - a for a, b in parts]
    gathered = shard.gather_frames(local, counts)             # ragged shards through one equal-sized all-gather
    if rank == 0:
        q.put((gathered, counts))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapping_partition_gather_and_merge_world2():
    """temporal_overlap > 0: ranks produce overlapping ranges (inference_cli.py:1166-1176), one all-gather, then the
    reference's merge with cross-fade (inference_cli.py:1241-1274) restores the frame count."""
    from oracle import color_oracle
    sys.path.insert(0, ROOT)
    from svr2_import import load_package
    load_package()
    import importlib
    shard = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.shard")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world, overlap = 11, 2, 3
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, 29612, total, overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert counts == [9, 5] and gathered.shape[0] == 14
    chunks = list(torch.split(gathered, counts, 0))
    for r, c in enumerate(chunks):                             # the gather returned every rank's frames unchanged
        assert torch.equal(c, torch.rand(counts[r], 4, 6, 3, generator=torch.Generator().manual_seed(100 + r)))
    blend = lambda p, c: color_oracle.blend_overlapping_frames(p, c, p.shape[0])
    out = shard.merge_shards(chunks, overlap, blend=blend)
    assert out.shape[0] == total
    assert torch.equal(out, color_oracle.merge_shards(chunks, overlap))
    assert torch.equal(out[:6], chunks[0][:6]) and torch.equal(out[9:], chunks[1][3:])
