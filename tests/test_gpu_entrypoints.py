"""train.py -> checkpoint -> resume -> export_arcflow_to_diffusers.py -> pipe.load_arcflow_adapter -> pipe(...): the
reference's user journey (train.sh / export.sh / inference_flux.py) on a depth-reduced FLUX with synthetic weights."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
CFG = os.path.join(ROOT, "configs/flux/arcflux_tiny_smoke.py")


def _train(tmp_path, total_iters):
    import train
    ck = str(tmp_path / "ck")
    return train.main([CFG, "--work-dir", str(tmp_path / "work"), "--seed", "11", "--cfg-options",
                       f"checkpoint_config.out_dir={ck}", f"resume_from={ck}/arcflux_tiny_smoke/latest.pth",
                       f"total_iters={total_iters}"])


def test_train_resume_export_load_infer(tmp_path):
    from lakonlab.runner import load_checkpoint
    runner = _train(tmp_path, 4)
    assert runner.iter == 4 and len(runner.log_buffer) == 4
    for lv in runner.log_buffer:
        assert lv["loss"] == lv["loss"] and lv["loss"] > 0          # finite
        assert "diffusion_grad_norm" in lv and "teacher_ratio" in lv
    ckdir = tmp_path / "ck" / "arcflux_tiny_smoke"
    assert (ckdir / "iter_2.pth").exists() and (ckdir / "iter_4.pth").exists() and (ckdir / "latest.pth").exists()
    ck = load_checkpoint(str(ckdir / "latest.pth"))
    assert ck["meta"]["iter"] == 4 and ck["meta"]["seed"] == 11
    keys = list(ck["state_dict"])
    assert any(k.startswith("diffusion_ema.denoising.") and "lora_A" in k for k in keys)
    assert all(v.dtype == torch.bfloat16 for v in ck["state_dict"].values())
    moved = ck["state_dict"]["diffusion.denoising.proj_out_logweights.weight"].float().abs().max().item()
    assert moved > 0                                                 # zero-initialised head was trained
    params4 = ck["optimizer"]["diffusion"]["params"].clone()
    del runner
    torch.cuda.empty_cache()

    # resume: picks up at iteration 4 with the saved arenas, runs to 6
    runner = _train(tmp_path, 6)
    assert runner.iter == 6 and len(runner.log_buffer) == 2
    assert (ckdir / "iter_6.pth").exists()
    ck6 = load_checkpoint(str(ckdir / "iter_6.pth"))
    assert ck6["optimizer"]["diffusion"]["steps_taken"] >= ck["optimizer"]["diffusion"]["steps_taken"]
    assert not torch.equal(ck6["optimizer"]["diffusion"]["params"], params4)

    # export the EMA adapter, load it through the inference surface, compare with the trainer's own engine
    import export_arcflow_to_diffusers as export
    out_dir = str(tmp_path / "adapter")
    export.main([CFG, str(ckdir / "iter_6.pth"), out_dir])
    from lakonlab.models.builder import student_config, synthetic_base_state_dict
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline, FluxBaseTransformer
    from lakonlab.utils import Config
    _, mc = student_config(dict(Config.fromfile(CFG).model.diffusion.denoising))
    pipe = ArcFluxPipeline(transformer=FluxBaseTransformer(synthetic_base_state_dict("flux", mc, 1234, "cuda"), device="cuda"))
    assert pipe.load_arcflow_adapter(out_dir) == "transformer_arcflow"
    g = torch.Generator().manual_seed(0)
    txt = (torch.randn(2, 64, 256, generator=g) * 0.1).to("cuda", torch.bfloat16)
    pooled = torch.randn(2, 256, generator=g).to("cuda", torch.bfloat16)
    lat = torch.randn(2, 64, 64, generator=g).to("cuda")
    got = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=lat.clone(), height=128, width=128,
               num_inference_steps=2, timestep_ratio=1.0, output_type="latent").images
    runner.model.trainer.write_back(use_ema=True)
    want = runner.model.diffusion.denoise(lat.clone(), txt, pooled, (8, 8), num_inference_steps=2, timestep_ratio=1.0)
    assert torch.isfinite(got).all() and torch.equal(got, want)
