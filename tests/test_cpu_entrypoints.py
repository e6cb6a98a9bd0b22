"""Host side of the train / inference entry points (no GPU): config system, adapter initialisation, datasets, the
iteration runner with its checkpoint / resume protocol, and the train.py command line."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lakonlab.utils import Config  # noqa: E402


def test_config_base_merge_and_overrides(tmp_path):
    (tmp_path / "base.py").write_text("opt = dict(lr=1e-4, betas=(0.9, 0.95))\nnested = dict(a=dict(x=1, y=2))\n")
    (tmp_path / "child.py").write_text("_base_ = ['./base.py']\nname = 'n'\nwork_dir = f'w/{name}'\n"
                                       "nested = dict(a=dict(y=3), b=4)\nrepl = dict(_delete_=True, z=1)\n")
    cfg = Config.fromfile(str(tmp_path / "child.py"))
    assert cfg.opt.lr == 1e-4 and cfg.nested.a.x == 1 and cfg.nested.a.y == 3 and cfg.nested.b == 4
    assert cfg.work_dir == "w/n" and cfg.repl == {"z": 1}
    cfg.merge_from_dict({"opt.lr": 5e-5, "new.k": [1, 2]})
    assert cfg.opt.lr == 5e-5 and cfg.new.k == [1, 2]
    assert "opt" in cfg.pretty_text and cfg.get("missing", 7) == 7


@pytest.mark.parametrize("path,student,layers", [("configs/flux/arcflux_2nfe_k16.py", "ArcFluxTransformer2DModel", 19),
                                                 ("configs/qwen/arcqwen_2nfe_k16.py", "ArcQwenImageTransformer2DModel", 60)])
def test_shipped_configs_carry_the_reference_hyperparameters(path, student, layers):
    """Values of configs/*/arc*_2nfe_k16.py + _ddp_train.py of the reference."""
    cfg = Config.fromfile(os.path.join(ROOT, path))
    d = cfg.model.diffusion.denoising
    assert d.type == student and d.num_layers == layers and d.num_gaussians == 16 and d.lora_rank == 256
    assert d.lora_dropout == 0.05 and d.logweights_channels == 4
    assert cfg.model.tie_teacher and cfg.model.diffusion.flow_loss.rescale_cfg.scale == 30.0
    assert cfg.model.diffusion.timestep_sampler.shift == 3.2
    t = cfg.train_cfg
    assert (t.num_decay_iters, t.window_substeps, t.gm_dropout, t.num_intermediate_states, t.nfe, t.total_substeps) == \
        (2000, 3, 0.1, 4, 2, 128)
    assert t.diffusion_grad_clip == 50.0 and t.diffusion_grad_clip_begin_iter == 100
    o = cfg.optimizer.diffusion
    assert o.lr == 1e-4 and tuple(o.betas) == (0.9, 0.95) and o.paramwise_cfg.custom_keys.proj_out_loggamma.lr_mult == 0.1
    assert cfg.lr_config.warmup_iters == 100 and cfg.lr_config.warmup_ratio == 0.001
    assert cfg.runner.type == "DynamicIterBasedRunnerMod" and cfg.runner.pass_training_status
    h = cfg.custom_hooks[0]
    assert h.type == "ExponentialMovingAverageHookMod" and h.start_iter == 100 and h.momentum_cfg.gamma == 7.0
    assert cfg.total_iters == 10000 and cfg.data.train_dataloader.samples_per_gpu == 4


def test_student_config_from_denoising_dict():
    from lakonlab.models.builder import student_config
    cfg = Config.fromfile(os.path.join(ROOT, "configs/flux/arcflux_tiny_smoke.py"))
    arch, mc = student_config(dict(cfg.model.diffusion.denoising))
    assert arch == "flux" and mc.num_layers == 2 and mc.num_single_layers == 2 and mc.lora_rank == 256 and mc.inner_dim == 256
    with pytest.raises(ValueError):
        student_config(dict(type="UNet"))
    from lakonlab.models.builder import load_transformer_weights
    with pytest.raises(ValueError, match="offline"):
        load_transformer_weights("huggingface://black-forest-labs/FLUX.1-dev/transformer/x.index.json")


def test_adapter_init_reproduces_the_teacher_velocity():
    """The reference initialises the student so that it starts AT the teacher (arcflux.py:92-132, :318-341): heads =
    proj_out repeated K times (+ a bias jitter shared by the 4 sub-pixels of a (component, channel)), log-weights 0,
    log-gamma weight 0 with the log-spaced bias, norm_out copied, LoRA B = 0. Checked through the oracle."""
    from arcflow_b200.adapter_init import flux_lora_target_paths, init_arcflow_adapter, loggamma_bias
    from arcflow_b200.config import FLUX_LORA_TARGETS, flux_tiny
    from arcflow_b200.synthetic import make_flux_inputs
    from lakonlab.models.builder import synthetic_base_state_dict
    from oracle import arcflow_oracle as O
    from oracle import arcflow_train_oracle as T
    cfg = flux_tiny()
    cfg.lora_rank = 8
    base = {k: v.float() for k, v in synthetic_base_state_dict("flux", cfg, 3, "cpu").items()}
    targets = flux_lora_target_paths(cfg, FLUX_LORA_TARGETS)
    assert len(targets) == 4 * cfg.num_layers + 2 * cfg.num_single_layers + 2
    ad = init_arcflow_adapter(base, cfg, targets, generator=torch.Generator().manual_seed(0), dtype=torch.float32)
    K, C, L = cfg.num_gaussians, cfg.out_channels, cfg.logweights_channels
    assert ad["proj_out_means.weight"].shape == (K * C, cfg.inner_dim)
    assert torch.equal(ad["proj_out_means.weight"].view(K, C, -1)[5], base["proj_out.weight"])
    jit = (ad["proj_out_means.bias"].view(K, C) - base["proj_out.bias"][None]).view(K, C // L, L)
    assert torch.allclose(jit, jit[..., :1].expand_as(jit), atol=1e-6) and 0.01 < jit.std() < 0.1
    assert not ad["proj_out_logweights.weight"].any() and not ad["proj_out_loggamma.weight"].any()
    g = ad["proj_out_loggamma.bias"].view(K - 1, L)
    assert torch.allclose(g[:, 0].exp(), torch.logspace(math.log10(0.2), math.log10(4.0), K - 1), rtol=1e-5)
    assert torch.equal(loggamma_bias(K, L), ad["proj_out_loggamma.bias"])
    assert all(not v.any() for k, v in ad.items() if "lora_B" in k)
    assert all(abs(v.std().item() - 1 / 8) < 0.03 for k, v in ad.items() if "lora_A" in k)
    sd = {k: v for k, v in base.items() if not k.startswith("proj_out.")}
    sd.update(ad)
    x, txt, pooled = make_flux_inputs(cfg, 1, 64, 64, txt_len=8, seed=1)
    sigma, gd = torch.tensor([0.7]), torch.tensor([3.5])
    out = O.flux_forward(sd, cfg, x, txt.float(), pooled.float(), sigma, gd, (4, 4))
    u = T.flux_teacher_velocity(base, cfg, x, txt.float(), pooled.float(), sigma, gd, (4, 4))
    mean_jit = jit[..., 0].reshape(K, C // L)
    for k in (0, 7, 15):
        want = u + mean_jit[k].repeat_interleave(L)[None, None]
        assert torch.allclose(out["means"][:, :, k], want, atol=2e-5)
    assert torch.allclose(out["logweights"].exp().sum(2), torch.ones(1, 16, L), atol=1e-5)
    assert torch.allclose(out["logweights"], torch.full_like(out["logweights"], -math.log(K)), atol=1e-5)


def test_datasets(tmp_path):
    from lakonlab.datasets import build_dataloader, build_dataset
    ds = build_dataset(dict(type="SyntheticPrompts", joint_attention_dim=32, pooled_projection_dim=8, seq_len=6,
                            latent_size=(16, 8, 8), length=10, negative=True))
    s = ds[3]
    assert s["prompt_embed_kwargs"]["encoder_hidden_states"].shape == (6, 32) and s["latents"].shape == (16, 8, 8)
    assert torch.equal(ds[3]["prompt_embed_kwargs"]["pooled_projections"], s["prompt_embed_kwargs"]["pooled_projections"])
    assert torch.equal(ds[0]["negative_prompt_embed_kwargs"]["encoder_hidden_states"],
                       ds[9]["negative_prompt_embed_kwargs"]["encoder_hidden_states"])
    batch = next(iter(build_dataloader(ds, samples_per_gpu=4)))
    assert batch["prompt_embed_kwargs"]["encoder_hidden_states"].shape == (4, 6, 32)
    # cached files: new-style and legacy keys (image_prompts.py:86-91), padded to pad_seq_len
    torch.save(dict(prompt_embed_kwargs=dict(encoder_hidden_states=torch.ones(3, 32).half(), pooled_projections=torch.ones(8))),
               tmp_path / "a.pt")
    torch.save(dict(prompt_embeds=torch.ones(9, 32), pooled_prompt_embeds=torch.ones(8)), tmp_path / "b.pt")
    ds2 = build_dataset(dict(type="ImagePrompts", cache_dir=str(tmp_path), pad_seq_len=5, latent_size=(16, 4, 4)))
    a, b = ds2[0]["prompt_embed_kwargs"], ds2[1]["prompt_embed_kwargs"]
    assert a["encoder_hidden_states"].shape == (5, 32) and a["encoder_hidden_states"].dtype == torch.float32
    assert a["encoder_hidden_states"][3:].abs().sum() == 0 and b["encoder_hidden_states"].shape == (5, 32)
    assert b["pooled_projections"].shape == (8,)


class _FakeOpt:
    def __init__(self):
        self.v = torch.zeros(3)

    def state_dict(self):
        return dict(v=self.v.clone())

    def load_state_dict(self, sd):
        self.v.copy_(sd["v"])


class _FakeModel:
    def __init__(self):
        self.opt = _FakeOpt()
        self.generator = torch.Generator().manual_seed(1)
        self.seen = []
        self.trainer = type("T", (), {"iteration": 0, "write_back": lambda self_: None, "opt": self.opt})()

    def train_step(self, data, optimizer, running_status=None):
        self.seen.append(running_status["iteration"])
        optimizer["diffusion"].v += 1
        return dict(log_vars=dict(loss=float(torch.rand(1, generator=self.generator))), num_samples=len(data["x"]))

    def state_dict(self, trainable_only=True):
        return {"diffusion.denoising.w": self.opt.v.clone(), "diffusion_ema.denoising.w": self.opt.v * 2}


def test_runner_loop_checkpoint_and_resume(tmp_path):
    from torch.utils.data import DataLoader
    from lakonlab.runner import CheckpointHook, DynamicIterBasedRunnerMod, adapter_from_checkpoint, load_checkpoint
    loader = DataLoader([dict(x=torch.zeros(1)) for _ in range(3)], batch_size=1)
    m = _FakeModel()
    r = DynamicIterBasedRunnerMod(m, optimizer={"diffusion": m.opt}, work_dir=str(tmp_path), max_iters=5,
                                  pass_training_status=True, gc_interval=2)
    r.register_hook(CheckpointHook(interval=2, out_dir=str(tmp_path), max_keep_ckpts=1))
    r.run([loader], [("train", 2)])
    assert m.seen == [0, 1, 2, 3, 4] and r.iter == 5 and r.epoch == 1
    assert not (tmp_path / "iter_2.pth").exists() and (tmp_path / "iter_4.pth").exists()   # max_keep_ckpts = 1
    ck = load_checkpoint(str(tmp_path / "latest.pth"))
    assert ck["meta"]["iter"] == 4 and torch.equal(ck["optimizer"]["diffusion"]["v"], torch.full((3,), 4.0))
    assert torch.equal(adapter_from_checkpoint(ck)["w"], torch.full((3,), 8.0))            # EMA preferred
    assert torch.equal(adapter_from_checkpoint(ck, use_ema=False)["w"], torch.full((3,), 4.0))
    m2 = _FakeModel()
    r2 = DynamicIterBasedRunnerMod(m2, optimizer={"diffusion": m2.opt}, work_dir=str(tmp_path), max_iters=6,
                                   pass_training_status=True)
    r2.resume(str(tmp_path / "latest.pth"))
    assert r2.iter == 4 and torch.equal(m2.opt.v, torch.full((3,), 4.0)) and m2.trainer.iteration == 4
    r2.run([loader], [("train", 10)])
    assert m2.seen == [4, 5] and r2.iter == 6
    with pytest.raises(NotImplementedError):
        DynamicIterBasedRunnerMod(m, ckpt_trainable_only=False)


def test_train_cli_flags_match_the_reference():
    """train.py:45-98 of the reference."""
    import train
    a = train.parse_args(["cfg.py", "--work-dir", "w", "--resume-from", "c.pth", "--no-validate", "--gpu-id", "1", "--seed", "7",
                          "--diff_seed", "--deterministic", "--cfg-options", "a.b=1", "name=x", "--launcher", "pytorch",
                          "--local_rank", "0"])
    assert a.config == "cfg.py" and a.work_dir == "w" and a.resume_from == "c.pth" and a.no_validate and a.gpu_id == 1
    assert a.seed == 7 and a.diff_seed and a.deterministic and a.cfg_options == {"a.b": 1, "name": "x"} and a.launcher == "pytorch"
    assert train.parse_args(["cfg.py"]).seed == 2021
