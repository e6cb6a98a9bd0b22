"""Pins oracle/vae_oracle.py on the original black-forest-labs autoencoder code shipped in this image
(torchtitan.experiments.flux.model.autoencoder — the module FLUX.1's VAE weights were trained with; diffusers'
AutoencoderKL, which the reference calls at arcflux_pipeline.py:531-534, is a conversion of it)."""
import pytest
import torch

from oracle import vae_oracle as V

A = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")


@pytest.mark.parametrize("ch,res,hw", [(32, 32, (4, 4)), (64, 64, (8, 6)), (32, 64, (5, 9))])
def test_decoder_restatement_matches_the_bfl_module(ch, res, hw):
    params = A.AutoEncoderParams(resolution=res, ch=ch)
    ae = A.AutoEncoder(params).float().eval()
    g = torch.Generator().manual_seed(ch + res)
    with torch.no_grad():
        for p in ae.parameters():                      # default inits give near-degenerate outputs; draw seeded values
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.3) + (1.0 if p.dim() == 1 else 0.0))
    sd = {k: v for k, v in ae.state_dict().items() if k.startswith("decoder.")}
    z = torch.randn(2, 16, *hw, generator=g)
    with torch.no_grad():
        want = ae.decode(z)
        got = V.vae_decode(sd, z, dtype=torch.float32, scale_factor=params.scale_factor, shift_factor=params.shift_factor)
        got64 = V.vae_decode({k: v.double() for k, v in sd.items()}, z.double(), dtype=torch.float64,
                             scale_factor=params.scale_factor, shift_factor=params.shift_factor)
    assert got.shape == want.shape == (2, 3, 8 * hw[0], 8 * hw[1])
    err = ((got - want).norm() / want.norm()).item()
    err64 = ((got64.float() - want).norm() / want.norm()).item()
    assert err < 1e-5 and err64 < 1e-4, (err, err64)


def test_synthetic_state_dict_has_the_bfl_layout():
    params = A.AutoEncoderParams()
    ref_keys = {k: tuple(v.shape) for k, v in A.AutoEncoder(params).state_dict().items() if k.startswith("decoder.")}
    sd = V.make_vae_decoder_state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == ref_keys
    assert V.SCALE_FACTOR == params.scale_factor and V.SHIFT_FACTOR == params.shift_factor
    small = V.make_vae_decoder_state_dict(ch=64, seed=3)
    z = torch.randn(1, 16, 4, 4, generator=torch.Generator().manual_seed(0))
    img = V.vae_decode(small, z, dtype=torch.float32)
    assert img.shape == (1, 3, 32, 32) and torch.isfinite(img).all() and 0.05 < img.std().item() < 50
