"""Pins oracle/arcflow_train_oracle.py (training roll-out: scheduled trajectory mixing, average-velocity matching,
mixture dropout, teacher Euler steps, loss and its gradients w.r.t. the policy parameters) against
tests/golden/reference_train_rollout.npz — produced by tools/make_golden_train.py executing the reference's own
ArcFlowImitationBase.piid_segment_momentum with a synthetic closed-form teacher."""
import numpy as np
import pytest
import torch

from oracle import arcflow_train_oracle as T
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def g():
    return np.load(ROOT / "tests" / "golden" / "reference_train_rollout.npz")


def teacher_u(x_t, t):
    return torch.tanh(x_t * 0.7) * (0.5 + t.reshape(-1, 1, 1, 1)) - 0.3 * x_t.flip(1)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_piid_segment_matches_reference(g, case):
    t = lambda k: torch.from_numpy(g[k])
    raw0, seg, ratio, p = [float(v) for v in g[f"c{case}_args"]]
    mp = dict(means=t("in_means").clone().requires_grad_(True), logweights=t("in_logw").clone().requires_grad_(True),
              loggammas=t("in_gam").clone().requires_grad_(True))
    x = t("in_x")
    B = x.shape[0]
    raw_t_src = torch.full((B,), raw0)
    sigma_src = T.warp_t(raw_t_src, 3.2).reshape(B, 1, 1, 1)
    rand = dict(drop_u=t(f"c{case}_drop_u"), student_u=t(f"c{case}_student_u"), teacher_u=t(f"c{case}_teacher_u"))
    cfg = dict(eps=1e-4, total_substeps=128, num_intermediate_states=4, window_substeps=3, gm_dropout=p)
    loss, x_dst, raw_dst, _ = T.piid_segment(mp, x, raw_t_src, sigma_src, ratio, seg, teacher_u, rand, cfg)
    loss.backward()
    assert float(loss) == pytest.approx(float(g[f"c{case}_loss"][0]), rel=1e-6)
    assert torch.allclose(x_dst, t(f"c{case}_x_dst"), atol=2e-6)
    assert torch.allclose(raw_dst, t(f"c{case}_raw_dst"), atol=1e-7)
    for name, key in (("means", "grad_means"), ("logweights", "grad_logw"), ("loggammas", "grad_gam")):
        ref = t(f"c{case}_{key}")
        assert torch.allclose(mp[name].grad, ref, atol=1e-7 + 1e-5 * ref.abs().max().item()), name


def test_dropout_mask_never_drops_everything():
    u = torch.tensor([[0.01] * 16, [0.5] * 15 + [0.01]])
    m = T.dropout_mask(u, 0.1)
    assert not m[0].any() and m[1].sum() == 1


def test_qwen_true_cfg_teacher_identities():
    """qwen_teacher_cfg_velocity: g <= 1 is the plain conditional velocity; neg == pos cancels the guidance term; the
    general case is pos + (pos - neg)(g - 1) of two independent bf16-rounded calls (gaussian_flow.py:18-26, 224-254)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from arcflow_b200.qwen import make_qwen_inputs, make_qwen_state_dict, make_qwen_teacher_extras, qwen_tiny
    from oracle import arcflow_train_oracle as T
    cfg = qwen_tiny(2, 2)
    sd, extra = make_qwen_state_dict(cfg, seed=3), make_qwen_teacher_extras(cfg, seed=4)
    x, txt = make_qwen_inputs(cfg, 2, 32, 32, txt_len=8, seed=5)
    _, neg = make_qwen_inputs(cfg, 2, 32, 32, txt_len=8, seed=6)
    tsd = T.teacher_state_dict(sd, extra)
    sig = torch.tensor([0.9, 0.3])
    plain = T.qwen_teacher_velocity(tsd, cfg, x.bfloat16(), txt, sig, (2, 2)).bfloat16().float()
    plain_neg = T.qwen_teacher_velocity(tsd, cfg, x.bfloat16(), neg, sig, (2, 2)).bfloat16().float()
    assert torch.equal(T.qwen_teacher_cfg_velocity(tsd, cfg, x.bfloat16(), txt, neg, sig, 1.0, (2, 2)), plain)
    assert torch.allclose(T.qwen_teacher_cfg_velocity(tsd, cfg, x.bfloat16(), txt, txt, sig, 4.0, (2, 2)), plain, atol=1e-6)
    got = T.qwen_teacher_cfg_velocity(tsd, cfg, x.bfloat16(), txt, neg, sig, 4.0, (2, 2))
    assert torch.allclose(got, plain + (plain - plain_neg) * 3.0, atol=2e-2, rtol=2e-2)   # batch-doubled vs separate calls
