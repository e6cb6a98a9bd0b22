"""GPU: the Qwen-Image distillation path — tied teacher with true CFG, train-step forward, and the adapter gradients —
against the training oracle (oracle/arcflow_train_oracle.py: qwen_teacher_cfg_velocity / qwen_train_forward)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import arcflow_train_oracle as T  # noqa: E402

DEV = "cuda"
TC = dict(num_decay_iters=2000, window_substeps=3, gm_dropout=0.1, num_intermediate_states=4, teacher_guidance_scale=4.0,
          nfe=2, timestep_ratio=1.0, total_substeps=128, eps=1e-4, lora_dropout=0.1)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def _setup(num_layers=3, heads=2, batch=2, px=64, txt_len=24):
    from arcflow_b200.qwen import (ArcQwenEngineModel, QwenTeacherEngine, make_qwen_inputs, make_qwen_state_dict,
                                   make_qwen_teacher_extras, qwen_tiny)
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = qwen_tiny(num_layers, heads)
    sd = make_qwen_state_dict(cfg, seed=4321, device="cpu")
    extra = make_qwen_teacher_extras(cfg, seed=7)
    x, txt = make_qwen_inputs(cfg, batch, px, px, txt_len=txt_len, seed=11, device="cpu")
    _, neg = make_qwen_inputs(cfg, batch, px, px, txt_len=txt_len, seed=12, device="cpu")
    student = ArcQwenEngineModel(sd, cfg, device=DEV)
    teacher = QwenTeacherEngine(student, extra)
    return cfg, sd, extra, x, txt, neg, (px // 16, px // 16), student, teacher


def test_qwen_tied_teacher_true_cfg_parity(lib, parity):
    cfg, sd, extra, x, txt, neg, grid, student, teacher = _setup()
    sig = [0.83, 0.41]
    tsd = T.teacher_state_dict(sd, extra)
    u = teacher.velocity(x.bfloat16().to(DEV), txt.to(DEV), neg.to(DEV), sig, 4.0, grid)
    assert u.dtype == torch.float32
    ref = T.qwen_teacher_cfg_velocity(tsd, cfg, x.bfloat16(), txt, neg, torch.tensor(sig), 4.0, grid, dtype=torch.float32)
    ref_bf = T.qwen_teacher_cfg_velocity(tsd, cfg, x.bfloat16(), txt, neg, torch.tensor(sig), 4.0, grid, dtype=torch.bfloat16)
    # same criterion as the student parity tests: no further from the fp32 oracle than the oracle's own bf16 run (x1.5)
    parity("qwen_tiny_teacher.cfg_velocity", rel(u, ref), max(1.5 * rel(ref_bf, ref), 2e-2))
    u1 = teacher.velocity(x.bfloat16().to(DEV), txt.to(DEV), None, sig, 1.0, grid)       # no guidance: plain bf16 velocity
    ref1 = T.qwen_teacher_velocity(tsd, cfg, x.bfloat16(), txt, torch.tensor(sig), grid, dtype=torch.float32)
    assert u1.dtype == torch.bfloat16
    parity("qwen_tiny_teacher.velocity", rel(u1, ref1), 2e-2)


@pytest.mark.parametrize("iteration", [0, 900])
def test_qwen_train_step_forward_loss_parity(lib, parity, iteration):
    from arcflow_b200.train import ArcFlowDistillStep, draw_rollout_randoms
    cfg, sd, extra, x, txt, neg, grid, student, teacher = _setup()
    g = torch.Generator().manual_seed(50 + iteration)
    rands = [draw_rollout_randoms(2, 4, 16, g) for _ in range(2)]
    step = ArcFlowDistillStep(student, teacher, TC)
    loss, lv, _ = step.forward(txt.to(DEV), None, grid, x.to(DEV), rands, iteration=iteration, neg_txt=neg.to(DEV))
    ref, ref_lv, _ = T.qwen_train_forward(sd, extra, cfg, txt, neg, grid, x, rands, iteration, TC, dtype=torch.float32)
    ref_bf, _, _ = T.qwen_train_forward(sd, extra, cfg, txt, neg, grid, x, rands, iteration, TC, dtype=torch.bfloat16)
    tol = max(1.5 * abs(float(ref_bf) - float(ref)), 2e-2 * abs(float(ref)))
    parity(f"qwen_tiny_train.it{iteration}.loss", abs(loss - float(ref)) / abs(float(ref)), tol / abs(float(ref)), floor=1e-4)
    assert lv["teacher_ratio"] == ref_lv["teacher_ratio"]


@pytest.mark.parametrize("stash", [False, True])
def test_qwen_adapter_gradients_match_autograd(lib, parity, stash):
    """forward_backward() on the Qwen student: every LoRA pair (img_mlp of all blocks, txt_mlp of blocks 0..L-2 — the last
    block's text tail is skipped in the backward as in the forward), the timestep embedder's pairs, heads and norm_out."""
    from arcflow_b200.train import ArcFlowDistillStep, draw_rollout_randoms
    cfg, sd, extra, x, txt, neg, grid, student, teacher = _setup()
    g = torch.Generator().manual_seed(91)
    rands = [draw_rollout_randoms(2, 4, 16, g) for _ in range(2)]
    step = ArcFlowDistillStep(student, teacher, TC)
    student.set_activation_stash(stash)   # False: per-block recompute; True: block outputs kept by the train forward
    loss, _, grads = step.forward_backward(txt.to(DEV), None, grid, x.to(DEV), rands, iteration=700, neg_txt=neg.to(DEV))
    names = student.trunk_lora_names() + list(student.embed_lora_shapes()) + ["proj_out_means.weight", "norm_out.linear.weight"]
    assert len(student.trunk_lora_names()) == 2 * (2 * cfg.num_layers + 2 * (cfg.num_layers - 1))
    ref_loss, _, ex = T.qwen_train_forward(sd, extra, cfg, txt, neg, grid, x, rands, 700, TC, dtype=torch.float32,
                                           require_grad=names)
    ref_loss.backward()
    for n in names:
        ref = ex["leaves"][n].grad
        got = grads[n].float().cpu()
        e = rel(got, ref)
        cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
        parity(f"qwen_tiny_grads.{'stash' if stash else 'recompute'}.{n}", e, 2e-2)
        assert cos > 0.9995, f"{n}: rel-L2 {e:.3e} cos {cos:.5f} (|ref| {ref.norm():.3e})"
