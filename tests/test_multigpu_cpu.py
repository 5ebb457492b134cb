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
