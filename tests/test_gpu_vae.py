"""GPU: the FLUX VAE decoder on the native kernels vs oracle/vae_oracle.py — which
tests/test_oracle_vae.py pins on the original BFL autoencoder code. Building blocks first (implicit-GEMM 3x3 convolution
incl. image edges / ragged tiles / narrow outputs / residual, GroupNorm + swish, upsampling, row softmax, layout kernels),
then whole decodes: width-reduced (ch 64) at several latent shapes, and the full-width decoder (ch 128: 512-channel mid
block, 16384-position attention at 1024 px) against the oracle on the CPU."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as V  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.fixture(scope="module")
def ops(lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from arcflow_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("n,h,w,cin,cout,res", [(1, 16, 16, 64, 256, False), (2, 8, 8, 64, 64, True), (1, 20, 37, 128, 128, True),
                                                (2, 33, 16, 64, 8, False), (1, 64, 48, 256, 512, True), (1, 5, 3, 64, 320, False)])
def test_conv3x3_implicit_gemm(ops, parity, n, h, w, cin, cout, res):
    g = torch.Generator(device=DEV).manual_seed(h * w + cin)
    x = torch.randn(n, h, w, cin, device=DEV, generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device=DEV, generator=g) * (1.0 / (9 * cin)) ** 0.5).bfloat16()
    b = (torch.randn(cout, device=DEV, generator=g) * 0.1).bfloat16()
    r = torch.randn(n, h, w, cout, device=DEV, generator=g).bfloat16() if res else None
    out = ops.conv3x3(x, ops.pack_conv3x3_weight(wt), b, res=r)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    assert out.shape == (n, h, w, cout) and torch.isfinite(out.float()).all()
    parity(f"vae_conv3x3.n{n}h{h}w{w}ci{cin}co{cout}{'r' if res else ''}", rel(out, ref), 4e-3)
    if res:   # in-place residual (out aliases res), and a channel-slice view as the output
        r2 = r.clone()
        ops.conv3x3(x, ops.pack_conv3x3_weight(wt), b, out=r2, res=r2)
        assert torch.equal(r2, out)
    wide = torch.zeros(n, h, w, cout + 64, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3(x, ops.pack_conv3x3_weight(wt), b, out=wide[..., 32:32 + cout], res=r)
    assert torch.equal(wide[..., 32:32 + cout], out) and wide[..., :32].abs().sum() == 0 and wide[..., 32 + cout:].abs().sum() == 0


def test_conv3x3_padded_channels_and_errors(ops):
    from arcflow_b200 import AfbError
    g = torch.Generator(device=DEV).manual_seed(1)
    wt = torch.randn(3, 16, 3, 3, device=DEV, generator=g).bfloat16() * 0.1          # 16 -> 3 channels, padded to 64 -> 8
    x16 = torch.randn(1, 12, 12, 16, device=DEV, generator=g).bfloat16()
    x = torch.zeros(1, 12, 12, 64, device=DEV, dtype=torch.bfloat16)
    x[..., :16] = x16
    out = ops.conv3x3(x, ops.pack_conv3x3_weight(wt, 64, 8))
    ref = F.conv2d(x16.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    assert rel(out[..., :3], ref) < 4e-3 and out[..., 3:].abs().sum() == 0
    with pytest.raises(AfbError, match="multiple of 64"):
        ops.conv3x3(x[..., :32].contiguous(), ops.pack_conv3x3_weight(wt, 32, 8))
    with pytest.raises(AfbError):
        ops.conv3x3(x.cpu(), ops.pack_conv3x3_weight(wt, 64, 8).cpu())


@pytest.mark.parametrize("n,h,w,c,silu", [(2, 8, 8, 64, True), (1, 37, 21, 128, True), (2, 64, 64, 256, False), (1, 16, 16, 512, True)])
def test_groupnorm_swish(ops, parity, n, h, w, c, silu):
    g = torch.Generator(device=DEV).manual_seed(c + h)
    x = (torch.randn(n, h, w, c, device=DEV, generator=g) * 2 + 0.7).bfloat16()
    ga = 1 + 0.2 * torch.randn(c, device=DEV, generator=g)
    be = 0.3 * torch.randn(c, device=DEV, generator=g)
    y = ops.groupnorm(x, ga, be, silu=silu)
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, ga, be, eps=1e-6)
    if silu:
        ref = ref * torch.sigmoid(ref)
    ref = ref.permute(0, 2, 3, 1)
    parity(f"vae_groupnorm.n{n}h{h}w{w}c{c}{'s' if silu else ''}", rel(y, ref), 4e-3)
    y2 = x.clone()
    ops.groupnorm(y2, ga, be, silu=silu, out=y2)          # in place
    assert torch.equal(y2, y)
    assert torch.equal(ops.groupnorm(x, ga, be, silu=silu), y)   # deterministic


def test_upsample_softmax_and_layout_kernels(ops):
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(2, 5, 7, 64, device=DEV, generator=g).bfloat16()
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), ref)
    s = (torch.randn(300, 1024, device=DEV, generator=g) * 4).bfloat16()
    want = torch.softmax(s.float(), -1)
    ops.softmax_rows_(s)
    assert rel(s, want) < 4e-3 and (s.float().sum(-1) - 1).abs().max() < 2e-2
    z = torch.randn(2, 16, 6, 10, device=DEV, generator=g)
    pre = ops.vae_pre(z, 64, 0.3611, 0.1159)
    assert pre.shape == (2, 6, 10, 64) and pre[..., 16:].abs().sum() == 0
    assert torch.equal(pre[..., :16], (z / 0.3611 + 0.1159).permute(0, 2, 3, 1).bfloat16())
    img = torch.randn(2, 6, 10, 8, device=DEV, generator=g).bfloat16()
    assert torch.equal(ops.vae_post(img, 3), img[..., :3].float().permute(0, 3, 1, 2))


@pytest.mark.parametrize("hw,batch", [((8, 8), 2), ((16, 8), 1), ((24, 16), 1)])
def test_vae_decode_width_reduced_vs_oracle(lib, parity, hw, batch):
    from arcflow_b200.vae import FluxVAEDecoder
    sd = V.make_vae_decoder_state_dict(ch=64, seed=5)
    z = torch.randn(batch, 16, *hw, generator=torch.Generator().manual_seed(hw[0])) * 0.8
    dec = FluxVAEDecoder(sd, device=DEV)
    img = dec.decode(z.to(DEV))
    ref = V.vae_decode(sd, z, dtype=torch.float32)
    ref_bf16 = V.vae_decode(sd, z, dtype=torch.bfloat16)
    assert img.shape == ref.shape == (batch, 3, 8 * hw[0], 8 * hw[1]) and img.dtype == torch.float32
    assert torch.isfinite(img).all()
    e, e16 = rel(img, ref), rel(ref_bf16, ref)
    parity(f"vae_decode_ch64.{hw[0]}x{hw[1]}b{batch}", e, max(3e-2, 1.5 * e16))
    assert torch.equal(dec.decode(z.to(DEV)), img)        # deterministic


def test_vae_decode_accepts_diffusers_keys(lib):
    """diffusers' AutoencoderKL names (what `pipe.vae.state_dict()` yields in the reference) load to the same decoder."""
    from arcflow_b200.vae import FluxVAEDecoder
    sd = V.make_vae_decoder_state_dict(ch=64, seed=6)
    levels = 4
    dsd = {}
    for k, v in sd.items():
        n = k.replace("decoder.mid.block_1.", "decoder.mid_block.resnets.0.").replace("decoder.mid.block_2.", "decoder.mid_block.resnets.1.")
        n = n.replace("decoder.mid.attn_1.norm.", "decoder.mid_block.attentions.0.group_norm.")
        for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
            n = n.replace(f"decoder.mid.attn_1.{a}.", f"decoder.mid_block.attentions.0.{b}.")
        if n.startswith("decoder.up."):
            parts = n.split(".")
            lvl = levels - 1 - int(parts[2])
            if parts[3] == "block":
                n = f"decoder.up_blocks.{lvl}.resnets.{parts[4]}." + ".".join(parts[5:]).replace("nin_shortcut", "conv_shortcut")
            else:
                n = f"decoder.up_blocks.{lvl}.upsamplers.0." + ".".join(parts[4:])
        n = n.replace("decoder.norm_out.", "decoder.conv_norm_out.")
        if "attentions.0.to_" in n and n.endswith(".weight"):
            v = v[:, :, 0, 0]
        dsd[n] = v
    z = torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(3))
    a = FluxVAEDecoder(sd, device=DEV).decode(z.to(DEV))
    b = FluxVAEDecoder(dsd, device=DEV).decode(z.to(DEV))
    assert torch.equal(a, b)


def test_vae_decode_full_width_vs_oracle(lib, parity):
    """The real decoder (ch 128, ch_mult (1, 2, 4, 4): 512-channel mid block and attention) on a 32 x 32 latent (256 px
    image, BASELINE.json configs[0] size) against the CPU oracle in fp32 and bf16."""
    from arcflow_b200.vae import FluxVAEDecoder
    sd = V.make_vae_decoder_state_dict(seed=7)
    z = torch.randn(1, 16, 32, 32, generator=torch.Generator().manual_seed(11)) * 0.8
    dec = FluxVAEDecoder(sd, device=DEV)
    img = dec.decode(z.to(DEV))
    with torch.no_grad():
        ref = V.vae_decode(sd, z, dtype=torch.float32)
        ref_bf16 = V.vae_decode(sd, z, dtype=torch.bfloat16)
    e, e16 = rel(img, ref), rel(ref_bf16, ref)
    parity("vae_decode_full.32x32b1", e, max(3e-2, 1.5 * e16))


def test_pipeline_decodes_with_the_native_vae(lib):
    """output_type='pt' through ArcFluxPipeline with a FluxVAEDecoder attached: latents -> unpack -> decode -> [0, 1] image."""
    from arcflow_b200.config import flux_tiny
    from arcflow_b200.model import ArcFluxEngineModel
    from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
    from arcflow_b200.vae import FluxVAEDecoder
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline
    cfg = flux_tiny(1, 1, 2)
    sd = make_flux_state_dict(cfg, seed=3, device="cpu")
    x, txt, pooled = make_flux_inputs(cfg, 2, 128, 128, txt_len=16, seed=4)
    vsd = V.make_vae_decoder_state_dict(ch=64, seed=8)
    pipe = ArcFluxPipeline(transformer=ArcFluxEngineModel(sd, cfg, device=DEV), vae=FluxVAEDecoder(vsd, device=DEV))
    kw = dict(prompt_embeds=txt, pooled_prompt_embeds=pooled, latents=x, height=128, width=128, num_inference_steps=2,
              timestep_ratio=1.0)
    lat = pipe(output_type="latent", **kw).images
    img = pipe(output_type="pt", **kw).images
    assert img.shape == (2, 3, 128, 128) and img.dtype == torch.float32 and 0.0 <= img.min() and img.max() <= 1.0
    want = V.vae_decode(vsd, pipe._unpack_latents(lat.cpu(), 128, 128, 8), dtype=torch.float32)
    assert rel(img, (want / 2 + 0.5).clamp(0, 1)) < 3e-2
    arr = pipe(output_type="np", **kw).images
    assert arr.shape == (2, 128, 128, 3)
