"""GPU: the reference-facing plugin surface — load_arcflow_adapter from the on-disk adapter format and
ArcFluxPipeline.__call__ — end to end through the C ABI, checked against the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import arcflow_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.fixture(scope="module")
def setup(lib, tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.synthetic import make_flux_state_dict, make_flux_inputs
    from lakonlab.pipelines.arcflow_loader import split_adapter_keys, write_adapter_folder
    cfg = flux_tiny(2, 2, 2)
    sd = make_flux_state_dict(cfg, seed=77, device="cpu")
    # split into "stock FLUX" base (has a proj_out, no heads/LoRA) and the ArcFlow adapter on disk
    adapter_keys = [k for k in sd if "lora" in k or k.startswith(("proj_out_", "norm_out."))]
    adapter = {k: sd[k] for k in adapter_keys}
    base = {k: v for k, v in sd.items() if k not in adapter or k.startswith("norm_out.")}
    base["norm_out.linear.weight"] = torch.zeros_like(sd["norm_out.linear.weight"])  # adapter must overwrite these
    base["norm_out.linear.bias"] = torch.zeros_like(sd["norm_out.linear.bias"])
    base["proj_out.weight"] = torch.zeros(64, cfg.inner_dim, dtype=torch.bfloat16)
    base["proj_out.bias"] = torch.zeros(64, dtype=torch.bfloat16)
    root = tmp_path_factory.mktemp("adapter")
    write_adapter_folder(root / "arcflow-flux-2steps", cfg, adapter)
    x, txt, pooled = make_flux_inputs(cfg, 2, 64, 64, txt_len=32, seed=5)
    return cfg, sd, base, root, x, txt, pooled


def test_load_arcflow_adapter_and_call(setup, parity):
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline, FluxBaseTransformer
    cfg, sd, base, root, x, txt, pooled = setup
    pipe = ArcFluxPipeline(transformer=FluxBaseTransformer(base, device="cuda"))
    name = pipe.load_arcflow_adapter(str(root), subfolder="arcflow-flux-2steps", target_module_name="transformer")
    assert name == "transformer_arcflow"
    assert pipe.transformer.num_gaussians == 16 and pipe.transformer.config.in_channels == 64
    out = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, height=64, width=64,
               num_inference_steps=2, timestep_ratio=1.0, output_type="latent").images
    ref = O.flux_denoise(sd, cfg, x, txt, pooled, (4, 4), num_inference_steps=2, timestep_ratio=1.0)
    assert out.shape == x.shape and out.dtype == torch.float32
    parity("pipeline_flux_tiny.latents", rel(out, ref), 2e-2)


def test_adapter_without_lora_returns_none(setup, tmp_path):
    from lakonlab.pipelines.arcflow_loader import write_adapter_folder
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline, FluxBaseTransformer
    cfg, sd, base, root, *_ = setup
    write_adapter_folder(tmp_path / "nolora", cfg, {"proj_out_means.bias": sd["proj_out_means.bias"]})
    pipe = ArcFluxPipeline(transformer=FluxBaseTransformer(base, device="cuda"))
    with pytest.warns(UserWarning, match="No LoRA"):
        assert pipe.load_arcflow_adapter(str(tmp_path / "nolora")) is None
    assert isinstance(pipe.transformer, FluxBaseTransformer)   # nothing swapped
    with pytest.raises(RuntimeError, match="load_arcflow_adapter"):
        pipe(prompt_embeds=torch.zeros(1, 8, 256), pooled_prompt_embeds=torch.zeros(1, 256), height=64, width=64)


def test_callback_path_equals_fused_loop(setup):
    """callback_on_step_end forces the per-step path (afb_engine_forward + afb_sampler_step); it must give the
    same latents as the single afb_engine_denoise call, and the callback sees every step."""
    from arcflow_b200.model import ArcFluxEngineModel
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline
    cfg, sd, base, root, x, txt, pooled = setup
    pipe = ArcFluxPipeline(transformer=ArcFluxEngineModel(sd, cfg, device="cuda"))
    kw = dict(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, height=64, width=64,
              num_inference_steps=4, timestep_ratio=0.5, output_type="latent")
    fused = pipe(**kw).images
    seen = []

    def cb(p, i, t, tensors):
        seen.append((i, float(t)))
        return {}

    stepped = pipe(callback_on_step_end=cb, **kw).images
    assert [i for i, _ in seen] == [0, 1, 2, 3] and seen[0][1] == pytest.approx(1000.0)
    assert torch.equal(fused, stepped)
    ref = O.flux_denoise(sd, cfg, x, txt, pooled, (4, 4), num_inference_steps=4, timestep_ratio=0.5)
    assert rel(fused, ref) < 2e-2


def test_generator_latents_and_errors(setup):
    from arcflow_b200.model import ArcFluxEngineModel
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline
    cfg, sd, base, root, x, txt, pooled = setup
    pipe = ArcFluxPipeline(transformer=ArcFluxEngineModel(sd, cfg, device="cuda"))
    kw = dict(prompt_embeds=txt[:1], pooled_prompt_embeds=pooled[:1], height=64, width=64, num_inference_steps=2,
              timestep_ratio=1.0, output_type="latent")
    a = pipe(generator=torch.Generator("cuda").manual_seed(42), **kw).images
    b = pipe(generator=torch.Generator("cuda").manual_seed(42), **kw).images
    assert a.shape == (1, 16, 64) and torch.equal(a, b) and torch.isfinite(a).all()
    with pytest.raises(NotImplementedError, match="text encoders"):
        pipe(prompt="a kangaroo", height=64, width=64)
    with pytest.raises(NotImplementedError, match="VAE"):
        pipe(**{**kw, "output_type": "pil"})
    with pytest.raises(ValueError, match="divisible by 16"):
        pipe(**{**kw, "height": 72})
    # reference signature kwargs outside the hot path are accepted by name and refused explicitly (arcflux_pipeline.py:268-269)
    with pytest.raises(NotImplementedError, match="IP-Adapter"):
        pipe(ip_adapter_image_embeds=[torch.zeros(1)], **kw)
    assert pipe(ip_adapter_image=None, **kw, generator=torch.Generator("cuda").manual_seed(42)).images.equal(a)
    assert pipe.to("cuda") is pipe and pipe.to(torch.bfloat16) is pipe
    with pytest.raises(RuntimeError, match="lives on"):
        pipe.to("cpu")


def test_adapter_keys_prefixed_with_the_target_module_name(setup, tmp_path):
    """Adapter files saved as `transformer.<key>` load like un-prefixed ones (reference arcflow_loader.py:246-250)."""
    from lakonlab.pipelines.arcflow_loader import read_adapter_folder, write_adapter_folder
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline, FluxBaseTransformer
    cfg, sd, base, root, x, txt, pooled = setup
    _, adapter = read_adapter_folder(str(root), "arcflow-flux-2steps")
    write_adapter_folder(tmp_path / "prefixed", cfg, {"transformer." + k: v for k, v in adapter.items()})
    outs = []
    for folder, sub in ((str(root), "arcflow-flux-2steps"), (str(tmp_path / "prefixed"), None)):
        pipe = ArcFluxPipeline(transformer=FluxBaseTransformer(dict(base), device="cuda"))
        assert pipe.load_arcflow_adapter(folder, subfolder=sub) == "transformer_arcflow"
        outs.append(pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, height=64, width=64,
                         num_inference_steps=2, timestep_ratio=1.0, output_type="latent").images)
    assert torch.equal(outs[0], outs[1])


def test_qwen_pipeline_adapter_roundtrip(lib, tmp_path):
    """ArcQwenImagePipeline + load_arcflow_adapter on the Qwen class name; mask trimming; parity vs oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from arcflow_b200.qwen import make_qwen_inputs, make_qwen_state_dict, qwen_tiny
    from lakonlab.pipelines.arcflow_loader import write_adapter_folder
    from lakonlab.pipelines.arcflux_pipeline import FluxBaseTransformer
    from lakonlab.pipelines.arcqwen_pipeline import ArcQwenImagePipeline
    cfg = qwen_tiny(2, 2)
    sd = make_qwen_state_dict(cfg, seed=21)
    adapter = {k: v for k, v in sd.items() if "lora" in k or k.startswith(("proj_out_", "norm_out."))}
    base = {k: v for k, v in sd.items() if k not in adapter}
    base["norm_out.linear.weight"] = torch.zeros_like(sd["norm_out.linear.weight"])
    base["norm_out.linear.bias"] = torch.zeros_like(sd["norm_out.linear.bias"])
    write_adapter_folder(tmp_path / "arcflow-qwen-2steps", cfg, adapter)
    pipe = ArcQwenImagePipeline(transformer=FluxBaseTransformer(base, device="cuda"))
    assert pipe.load_arcflow_adapter(str(tmp_path), subfolder="arcflow-qwen-2steps") == "transformer_arcflow"
    x, txt = make_qwen_inputs(cfg, 2, 64, 64, txt_len=48, seed=3)
    mask = torch.zeros(2, 48, dtype=torch.long)
    mask[0, :30] = 1
    mask[1, :24] = 1            # longest prompt: 30 tokens -> trimmed to 30
    out = pipe(prompt_embeds=txt, prompt_embeds_mask=mask, latents=x, height=64, width=64, num_inference_steps=2,
               timestep_ratio=1.0, output_type="latent").images
    ref = O.qwen_denoise(sd, cfg, x, txt[:, :30], (4, 4), num_inference_steps=2)
    assert rel(out, ref) < 2e-2
    # captured loop = eager loop, bit for bit, also when replayed on new inputs of the same shape; fused adapter runs
    pipe.enable_cuda_graph()
    g1 = pipe(prompt_embeds=txt, prompt_embeds_mask=mask, latents=x, height=64, width=64, num_inference_steps=2,
              timestep_ratio=1.0, output_type="latent").images
    eager2 = pipe.enable_cuda_graph(False)(prompt_embeds=txt * 0.5, prompt_embeds_mask=mask, latents=x.flip(0), height=64,
                                            width=64, num_inference_steps=2, timestep_ratio=1.0, output_type="latent").images
    g2 = pipe.enable_cuda_graph()(prompt_embeds=txt * 0.5, prompt_embeds_mask=mask, latents=x.flip(0), height=64, width=64,
                                  num_inference_steps=2, timestep_ratio=1.0, output_type="latent").images
    assert torch.equal(g1, out) and torch.equal(g2, eager2)
    fused = pipe.enable_cuda_graph(False).fuse_lora()(prompt_embeds=txt, prompt_embeds_mask=mask, latents=x, height=64,
                                                      width=64, num_inference_steps=2, timestep_ratio=1.0,
                                                      output_type="latent").images
    assert 0 < rel(fused, out) < 2e-2


def test_cuda_graph_replay_equals_eager_loop(setup):
    """pipe.enable_cuda_graph(): the captured denoising loop is bit-identical to the eager launch sequence, also when
    replayed with new inputs of the same shape, and a different schedule gets its own graph."""
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline, FluxBaseTransformer
    from arcflow_b200 import _lib
    cfg, sd, base, root, x, txt, pooled = setup
    pipe = ArcFluxPipeline(transformer=FluxBaseTransformer(base, device="cuda"))
    pipe.load_arcflow_adapter(str(root), subfolder="arcflow-flux-2steps")
    kw = dict(height=64, width=64, timestep_ratio=1.0, output_type="latent")
    eager = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, num_inference_steps=2, **kw).images
    eager2 = pipe(prompt_embeds=txt * 0.5, pooled_prompt_embeds=pooled, latents=x.flip(0), num_inference_steps=2, **kw).images
    eager4 = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, num_inference_steps=4, **kw).images
    pipe.enable_cuda_graph()
    g1 = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, num_inference_steps=2, **kw).images
    lib_ = _lib.load()
    n0 = lib_.afb_launch_count()
    g2 = pipe(prompt_embeds=txt * 0.5, pooled_prompt_embeds=pooled, latents=x.flip(0), num_inference_steps=2, **kw).images
    assert lib_.afb_launch_count() == n0          # replay: no kernel goes through the launchers again
    g4 = pipe(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, num_inference_steps=4, **kw).images
    assert torch.equal(g1, eager) and torch.equal(g2, eager2) and torch.equal(g4, eager4)
    assert len(pipe.transformer._graphs) == 2


def test_fuse_lora_equals_an_engine_built_from_merged_weights(lib):
    """`fuse_lora()` (W <- W + B A on the GEMM kernel, LoRA branches switched off) against a fresh engine whose state dict
    was merged on the host and carries no adapter branches; and against the un-merged path (bf16-level difference only)."""
    import dataclasses
    from arcflow_b200 import AfbError
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.model import ArcFluxEngineModel
    from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
    cfg = flux_tiny(2, 2, 2)
    sd = make_flux_state_dict(cfg, seed=1234, device="cpu")
    x, txt, pooled = make_flux_inputs(cfg, 2, 64, 64, txt_len=64, seed=42)
    args = (x.cuda(), txt.cuda(), pooled.cuda(), (4, 4))
    model = ArcFluxEngineModel(sd, cfg, device="cuda")
    unfused = model.denoise(args[0].clone(), *args[1:], num_inference_steps=2)
    model.fuse_lora()
    fused = model.denoise(args[0].clone(), *args[1:], num_inference_steps=2)
    merged = {}
    for k, v in sd.items():
        if "lora" in k:
            continue
        pre = k[:-len(".weight")]
        if k.endswith(".weight") and pre + ".lora_A.weight" in sd:
            v = (v.float() + sd[pre + ".lora_B.weight"].float() @ sd[pre + ".lora_A.weight"].float()).bfloat16()
        merged[k] = v
    plain = ArcFluxEngineModel(merged, dataclasses.replace(cfg, lora_rank=0), device="cuda")
    want = plain.denoise(args[0].clone(), *args[1:], num_inference_steps=2)

    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()
    assert rel(fused, want) < 2e-3, rel(fused, want)          # same merged bf16 weights up to fp32 summation order
    assert 0 < rel(fused, unfused) < 2e-2, rel(fused, unfused)  # one extra bf16 rounding of the merged weights
    with pytest.raises(AfbError):
        model.set_lora_scale(0.5)
    with pytest.raises(AfbError):
        model.forward_heads(*args[:3], 0.7, 3.5, (4, 4), train=True)
