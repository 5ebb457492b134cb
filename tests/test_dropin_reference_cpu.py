"""CPU (build container only): the engine's model-slot modules dropped into the REFERENCE's own pipeline objects.

The reference's ``VideoDiffusionInfer`` (``src/core/infer.py``) and its memory manager (``src/optimization/memory_manager.py``)
are imported from ``/root/reference`` through the test-only stubs of ``oracle/ref_import.py`` (+ a dict-backed
``omegaconf`` shim); ``runner.dit`` / ``runner.vae`` are the engine's ``B200NaDiT`` / ``B200VideoVAE``.  There is no GPU
here, so the engine's kernel calls are monkeypatched: the forwards delegate to the (reference-pinned) oracle on the same
weights.  What is under test is everything BETWEEN the reference and the kernels: the call contract of the slots
(``infer.py:117-199, 203-278, 315-395``: argument names, shapes, dtypes, the ``tiled`` / ``tile_size`` / ``tile_overlap``
keywords, ``.latent`` / ``.sample`` / ``.vid_sample``), the ``nn.Module`` surface the pipeline relies on
(``parameters()`` sniffing, ``named_modules()``, ``.to()``, ``requires_grad_().eval()``) and the lifecycle
(``manage_model_device``, ``clear_rope_lru_caches``, ``cleanup_dit`` / ``release_model_memory``,
``memory_manager.py:427-455, 544-581, 670-738, 1011-1097``).  ``/root/reference`` does not exist on the GPU box: skipped there.
"""
import importlib
import os
import sys
import types

import pytest
import torch
import yaml

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")


class Cfg(dict):
    """dict-backed stand-in for omegaconf.DictConfig: attribute access, .get, nested."""

    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = Cfg(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class ListCfg(list):
    pass


class Debug:
    def log(self, *a, **k):
        pass

    def start_timer(self, *a, **k):
        pass

    def end_timer(self, *a, **k):
        return 0.0

    def log_memory_state(self, *a, **k):
        pass


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_import
    ref_import.install_stubs()
    om = types.ModuleType("omegaconf")
    om.DictConfig, om.ListConfig = Cfg, ListCfg
    om.OmegaConf = types.SimpleNamespace(load=lambda p: Cfg(yaml.safe_load(open(p))), create=lambda x: Cfg(x),
                                         register_new_resolver=lambda *a, **k: None)
    sys.modules.setdefault("omegaconf", om)
    infer = importlib.import_module("src.core.infer")
    mm = importlib.import_module("src.optimization.memory_manager")
    cfg = Cfg(yaml.safe_load(open(os.path.join(REF, "configs_3b", "main.yaml"))))
    cfg.vae.dtype = "bfloat16"
    cfg.diffusion.cfg.scale = 1.0                      # generation_phases.py:598-601: one-step, cfg 1
    cfg.diffusion.timesteps.sampling.steps = 1
    return types.SimpleNamespace(infer=infer, mm=mm, cfg=cfg)


@pytest.fixture()
def engines(pkg, monkeypatch):
    """B200NaDiT / B200VideoVAE built on the CPU with the kernel layer replaced by the oracle (test double)."""
    from oracle import dit_oracle, vae_oracle
    lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
    dit = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.dit")
    vae = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.vae")
    monkeypatch.setattr(lib, "device_check", lambda: (148, 10, 0))

    def fake_linear(a, w, *, bias=None, epi=0, **kw):          # the three load-time time-embedding GEMMs
        y = a.float() @ w.float().T + (bias.float() if bias is not None else 0)
        y = y.to(torch.bfloat16)
        return torch.nn.functional.silu(y.float()).to(torch.bfloat16) if epi & lib.EPI_SILU else y
    monkeypatch.setattr(lib, "linear", fake_linear)
    over = dict(dim=256, heads=2, layers=4, mm_layers=2, txt_in_dim=64)
    cfg = dit.dit_config("3b", **over)
    dsd = pkg.weights.synth_dit_state_dict(cfg, seed=1234, dtype=torch.float16)
    vsd = pkg.weights.synth_vae_state_dict(seed=4321, dtype=torch.float16)
    d, v = dit.B200NaDiT(cfg, dsd, device="cpu"), vae.B200VideoVAE(vsd, device="cpu")
    calls = {"dit": [], "enc": [], "dec": []}
    ocfg = dit_oracle.dit_config("3b", **over)
    d32, v32 = {k: t.float() for k, t in dsd.items()}, {k: t.float() for k, t in vsd.items()}

    def dit_forward(self, vid, txt, vid_shape, txt_shape, timestep=None, disable_cache=False):
        calls["dit"].append(dict(vid=tuple(vid.shape), txt=tuple(txt.shape), vid_shape=vid_shape.tolist(),
                                 txt_shape=txt_shape.tolist(), timestep=timestep.tolist(), dtype=vid.dtype))
        (T, H, W), = vid_shape.tolist()
        return dit.NaDiTOutput(dit_oracle.dit_forward(d32, ocfg, vid.float(), txt.float(), T, H, W).to(vid.dtype))

    def enc(self, x, return_dict=True, tiled=False, tile_size=None, tile_overlap=None):
        calls["enc"].append(dict(shape=tuple(x.shape), tiled=tiled, tile_size=tile_size, tile_overlap=tile_overlap))
        return vae.VAEOutput(latent=vae_oracle.vae_encode(v32, x.float()).to(x.dtype).squeeze(2), latent_dist=None)

    def dec(self, z, return_dict=True, tiled=False, tile_size=None, tile_overlap=None):
        calls["dec"].append(dict(shape=tuple(z.shape), tiled=tiled, tile_size=tile_size, tile_overlap=tile_overlap))
        z5 = z.unsqueeze(2) if z.ndim == 4 else z
        return vae.VAEOutput(sample=vae_oracle.vae_decode(v32, z5.float()).to(z.dtype).squeeze(2))
    monkeypatch.setattr(dit.B200NaDiT, "forward", dit_forward)
    monkeypatch.setattr(vae.B200VideoVAE, "encode", enc)
    monkeypatch.setattr(vae.B200VideoVAE, "decode", dec)
    return types.SimpleNamespace(dit=d, vae=v, calls=calls, d32=d32, v32=v32, ocfg=ocfg)


def test_reference_runner_drives_the_engine_slots(ref, engines):
    """One clip through the reference's own VideoDiffusionInfer with the engine modules in its slots."""
    from oracle import dit_oracle, vae_oracle
    runner = ref.infer.VideoDiffusionInfer(ref.cfg, Debug(), encode_tiled=False, decode_tiled=True,
                                           decode_tile_size=(64, 64), decode_tile_overlap=(16, 16))
    runner.dit, runner.vae = engines.dit, engines.vae
    runner.configure_diffusion(device=torch.device("cpu"), dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(3)
    clip = (torch.rand(3, 5, 32, 48, generator=g) * 2 - 1).to(torch.bfloat16)            # c t h w (generation_phases.py:489)
    lat, = runner.vae_encode([clip])
    assert tuple(lat.shape) == (2, 4, 6, 16) and lat.dtype == torch.bfloat16              # t h w c, scaled by 0.9152
    want = vae_oracle.runner_encode(engines.v32, clip[None].float())
    assert (lat.float() - want).abs().max() < 0.05
    noise = torch.randn(lat.shape, generator=g).to(torch.bfloat16)
    cond = runner.get_condition(noise, task="sr", latent_blur=lat)
    txt = torch.randn(58, 64, generator=g).to(torch.bfloat16)
    x0, = runner.inference(noises=[noise], conditions=[cond], texts_pos=[txt], texts_neg=[txt])
    call, = engines.calls["dit"]
    assert call["vid"] == (2 * 4 * 6, 33) and call["txt"] == (58, 64) and call["vid_shape"] == [[2, 4, 6]]
    assert call["txt_shape"] == [[58]] and call["timestep"] == [1000.0]                    # the t the engine folds at load
    vid = torch.cat([noise, cond], -1).reshape(-1, 33).float()
    v = dit_oracle.dit_forward(engines.d32, engines.ocfg, vid, txt.float(), 2, 4, 6)
    assert (x0.float() - dit_oracle.one_step_latent(noise.float(), v.view(2, 4, 6, 16))).abs().max() < 0.1
    out, = runner.vae_decode([x0])
    assert tuple(out.shape) == (3, 5, 32, 48)
    assert engines.calls["enc"][0]["tiled"] is False
    assert engines.calls["dec"][0] == dict(shape=(1, 16, 2, 4, 6), tiled=True, tile_size=(64, 64), tile_overlap=(16, 16))


def test_module_surface_and_lifecycle(ref, engines):
    mm = ref.mm
    d, v = engines.dit, engines.vae
    for m in (d, v):
        assert isinstance(m, torch.nn.Module)
        p = next(m.parameters())                                  # device / dtype sniffing (generation_phases.py:620,708)
        assert p.device.type == "cpu" and p.dtype == torch.bfloat16 and not p.requires_grad
        assert len(list(m.buffers())) > 10
        m.requires_grad_(False).eval()                            # model_configuration.py:1240-1245
        assert m.to(torch.float16) is m and next(m.buffers()).dtype != torch.float16     # dtype casts are refused
        assert m.half() is m and m.float() is m
    # the attention seam is found the way apply_model_specific_config finds it (model_configuration.py:1206-1210)
    hits = [mod for mod in d.modules() if type(mod).__name__ == "FlashAttentionVarlen"]
    assert len(hits) == 1
    hits[0].attention_mode, hits[0].compute_dtype = "sdpa", torch.bfloat16
    assert mm.clear_rope_lru_caches(d) == 0                       # walks named_modules() without tripping
    n_buf = sum(b.numel() for b in d.buffers())
    mm.manage_model_device(model=d, target_device=torch.device("cpu"), model_name="DiT", debug=Debug(), reason="test")
    assert sum(b.numel() for b in d.buffers()) == n_buf
    # offloaded / released weights must fail loudly, never fall back
    lib = importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
    with pytest.raises(lib.Svr2Error):
        d._require_cuda("forward")
    runner = types.SimpleNamespace(dit=d, vae=v, sampler=1, schedule=1, sampling_timesteps=1)
    mm.cleanup_dit(runner, debug=Debug(), cache_model=False)      # memory_manager.py:1011-1097
    assert runner.dit is None and runner.sampler is None
    mm.cleanup_vae(runner, debug=Debug(), cache_model=False)
    assert runner.vae is None
