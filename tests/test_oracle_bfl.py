"""The oracle's FLUX block arithmetic against an independent implementation: the original black-forest-labs FLUX model
code as shipped in torchtitan (tools/make_golden_bfl.py generated tests/golden/bfl_flux_tiny.npz with it)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import make_golden_bfl as G  # noqa: E402
from oracle import arcflow_train_oracle as T  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "bfl_flux_tiny.npz")


def _oracle_out(cfg, sd, img, txt, y, t, grid, dtype):
    return T.flux_teacher_velocity(sd, cfg, img, txt, y, t, None, grid, dtype=dtype, bf16_quirks=False)


def test_oracle_matches_bfl_flux_golden():
    g = np.load(GOLDEN)
    cfg = G.tiny_cfg()
    sd = G.diffusers_state_dict(cfg)
    probe = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    assert probe == pytest.approx(float(g["weight_abs_sum"][0]), rel=1e-12), "seeded weights differ from the generator's"
    img, txt, y, t = (torch.from_numpy(g[k]) for k in ("img", "txt", "y", "t"))
    gi, ti, yi, tt = G.make_inputs(cfg, img.shape[0], txt.shape[1], int(g["grid"][0]), int(g["grid"][1]))
    assert torch.equal(gi, img) and torch.equal(ti, txt) and torch.equal(yi, y) and torch.equal(tt, t)
    ref = torch.from_numpy(g["out"])
    out = _oracle_out(cfg, sd, img, txt, y, t, tuple(int(v) for v in g["grid"]), torch.float32)
    err = (out - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item() + 1e-6, err
    out64 = _oracle_out(cfg, {k: v.double() for k, v in sd.items()}, img.double(), txt.double(), y.double(), t.double(),
                        tuple(int(v) for v in g["grid"]), torch.float64)
    assert (out64.float() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6


@pytest.mark.skipif(importlib.util.find_spec("torchtitan") is None, reason="torchtitan (BFL FLUX code) not in this image")
@pytest.mark.parametrize("gh,gw,st,seed", [(2, 3, 5, 11), (5, 4, 16, 12)])
def test_oracle_matches_bfl_flux_live(gh, gw, st, seed):
    """Same comparison on other shapes / seeds, running the BFL code in-process when it is importable."""
    cfg = G.tiny_cfg()
    sd = G.diffusers_state_dict(cfg, seed=seed)
    img, txt, y, t = G.make_inputs(cfg, 1, st, gh, gw, seed=seed)
    ref = G.run_bfl(sd, cfg, img, txt, y, t, gh, gw)
    out = _oracle_out(cfg, sd, img, txt, y, t, (gh, gw), torch.float32)
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
