"""Pins oracle/arcflow_oracle.py (sampler side) against vectors produced by the reference's own code
(tools/make_golden.py executed lakonlab/pipelines/arcflux_pipeline.py and
lakonlab/models/diffusions/policies/arcflow.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import arcflow_oracle as O

K = 16


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("nfe", [1, 2, 4, 8])
@pytest.mark.parametrize("ratio", [1.0, 0.5])
def test_schedule_matches_reference(golden, nfe, ratio):
    raw, sub, tot = O.retrieve_raw_timesteps(nfe, 128, ratio)
    key = f"sched_nfe{nfe}_r{int(ratio * 10)}"
    assert np.array_equal(np.asarray(raw, dtype=np.float64), golden[key + "_raw"])
    assert list(sub) == golden[key + "_sub"].tolist()
    assert tot == int(golden[key + "_tot"][0])


def test_schedule_golden_values():
    # nfe=2, ratio 1.0, shift 3.2 -> sigma 1.0, 0.761905, 0
    raw, sub, tot = O.retrieve_raw_timesteps(2, 128, 1.0)
    ts = O.scheduler_timesteps(raw, 3.2)
    assert sub == [64, 64] and tot == 128
    assert ts[0].item() == pytest.approx(1000.0, abs=1e-3)
    assert ts[64].item() == pytest.approx(761.9048, abs=1e-3)


def _mp(golden):
    return dict(means=_t(golden["in_means_tok"]), logweights=_t(golden["in_logw_tok"]),
                loggammas=_t(golden["in_gam_tok"]))


def test_layout_matches_reference(golden):
    mp = O.unpack_mp(_mp(golden), 4, 4, K)
    assert torch.equal(mp["means"], _t(golden["unpacked_means"]))
    assert torch.equal(mp["logweights"], _t(golden["unpacked_logw"]))
    assert torch.equal(mp["loggammas"], _t(golden["unpacked_gam"]))
    x_img = O.unpack_latents(_t(golden["in_x_tok"]), 4, 4)
    assert torch.equal(x_img, _t(golden["unpacked_x"]))
    assert torch.equal(O.pack_latents(x_img), _t(golden["repacked_x"]))
    assert torch.equal(O.pack_latents(x_img), _t(golden["in_x_tok"]))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_momentum_integration_matches_reference(golden, case):
    mp = O.unpack_mp(_mp(golden), 4, 4, K)
    x_img = O.unpack_latents(_t(golden["in_x_tok"]), 4, 4)
    s_src, s_start, s_end = [float(np.float32(v)) for v in golden[f"mi{case}_args"]]
    # the reference divides raw_t_end (fp32) by 1000 in fp32
    s_end = float((torch.tensor(golden[f"mi{case}_args"][2] * 1000.0, dtype=torch.float32) / 1000.0).item())
    x_end = O.momentum_integration(mp, x_img, s_src, s_start, s_end, eps=1e-4)
    ref = _t(golden[f"mi{case}_x_end"])
    assert torch.allclose(x_end, ref, rtol=0, atol=2e-6), (x_end - ref).abs().max()
    assert torch.allclose(O.pack_latents(x_end), _t(golden[f"mi{case}_x_end_tok"]), rtol=0, atol=2e-6)
    vel = O.policy_velocity(mp, s_src, s_start)
    assert torch.allclose(vel, _t(golden[f"mi{case}_velocity"]), rtol=0, atol=2e-6)


def test_integration_closed_forms():
    # lambda == 0 for every component -> Euler step with the mixture-mean velocity; K == 1 -> exact Euler
    g = torch.Generator().manual_seed(0)
    B, C, H, W = 2, 16, 4, 4
    means = torch.randn(B, K, C, H, W, generator=g, dtype=torch.float64)
    logw = torch.randn(B, K, 1, H, W, generator=g, dtype=torch.float64)
    gam = torch.zeros(B, K - 1, 1, H, W, dtype=torch.float64)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    out = O.momentum_integration(dict(means=means, logweights=logw, loggammas=gam), x, 1.0, 1.0, 0.4)
    ref = x - 0.6 * (torch.softmax(logw, 1) * means).sum(1)
    # phi is evaluated at the clamped |z| = 1e-4 -> (e^z - 1)/z = 1 + z/2 + ...
    assert torch.allclose(out, ref, atol=1e-3)
    one = dict(means=means[:, :1], logweights=logw[:, :1], loggammas=gam[:, :0])
    assert torch.allclose(O.momentum_integration(one, x, 1.0, 1.0, 0.4), x - 0.6 * means[:, 0], atol=1e-12)
