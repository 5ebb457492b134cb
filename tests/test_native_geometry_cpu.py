"""CPU: the native host runtime's geometry (csrc/engine.cu: window boxes, layouts, RoPE tables — what svr2_dit_forward builds
per clip shape) against the Python module's (dit.py, itself checked against the oracle / the reference's window.py for
the five BASELINE geometries).  engine.cu is compiled with SVR2_HOST_TEST (tables kept in host memory) by nvcc's host
compiler; skipped where nvcc is missing."""
import importlib
import math
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("geo") / "geometry_dump")
    csrc = os.path.join(ROOT, "comfyui-seedvr2_videoupscaler_b200", "csrc")
    r = subprocess.run([NVCC, "-std=c++17", "-O1", "-I", csrc, "-o", exe, os.path.join(ROOT, "tests", "native", "geometry_dump.cu"),
                        "-Xlinker", "--unresolved-symbols=ignore-all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def _freqs(variant, dtype):
    if variant == "3b":
        return (1.0 / (10000 ** (torch.arange(0, 42, 2)[:21].float() / 42))).to(dtype)
    return (torch.linspace(1.0, 128.0, 10) * math.pi).to(dtype)


@pytest.mark.parametrize("variant,geom,dtype", [
    ("3b", (1, 32, 32), torch.float16), ("3b", (5, 68, 120), torch.float16), ("3b", (3, 135, 240), torch.float16),
    ("3b", (17, 135, 240), torch.float16), ("3b", (3, 20, 36), torch.bfloat16), ("3b", (2, 17, 23), torch.float32),
    ("7b", (2, 135, 240), torch.float16), ("7b", (3, 20, 36), torch.float16), ("7b", (5, 33, 47), torch.float16),
])
def test_native_geometry_matches_python(pkg, dumper, variant, geom, dtype):
    dit = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.dit")
    T, Hp, Wp = geom
    l = 58
    fr = _freqs(variant, dtype)
    dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype]
    out = subprocess.run([dumper, str(T), str(Hp), str(Wp), str(l), str(int(variant == "7b")), str(dt), str(fr.numel())]
                         + [repr(float(x)) for x in fr], capture_output=True, text=True, check=True).stdout.split("\n")
    it = iter(out)
    for s, shifted in ((0, False), (1, True)):
        lay, size_rows = dit.build_layout(T, Hp, Wp, l, shifted, variant, "cpu")
        hdr = next(it).split()
        assert hdr[0] == "layout" and [int(x) for x in hdr[2:]] == [lay.n_win, lay.total, lay.max_len, lay.txt_rows.numel()]
        for name, want in (("cu", lay.cu_seqlens), ("row_src", lay.row_src), ("row_rope", lay.row_rope), ("out_row_map", lay.out_row_map),
                           ("tok_dst", lay.tok_dst), ("tok_rope", lay.tok_rope), ("txt_rows", lay.txt_rows)):
            line = next(it).split()
            assert line[0] == name and [int(x) for x in line[1:]] == want.reshape(-1).tolist(), f"{name} (shifted={shifted})"
        c, sn = dit.rope_tables(fr, variant, int(lay.row_rope.max().item()) + 1, size_rows)
        rows = int(next(it).split()[1])
        assert rows == c.shape[0]
        got = torch.tensor([[float(x) for x in next(it).split()] for _ in range(rows * fr.numel())])
        gc, gs = got[:, 0].view_as(c), got[:, 1].view_as(sn)
        # cos / sin come from libm here and from torch's vectorised kernels in dit.py: identical except where a value sits
        # within an fp32 ulp of a rounding tie of the table dtype (e.g. sin(300.0) in fp16) — at most a handful of entries
        # (fp32 tables: the two libraries differ by an fp32 ulp in a few per cent of the entries)
        for a, b in ((gc, c), (gs, sn)):
            diff = (a - b).abs()
            if dtype == torch.float32:
                assert diff.max().item() <= 2.4e-7, diff.max().item()
            else:
                assert (diff > 0).sum().item() <= max(2, a.numel() // 2000) and diff.max().item() <= 2 ** -9, (diff > 0).sum().item()
