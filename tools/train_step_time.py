"""Times the native train-step FORWARD (2 student + 8 teacher forwards + roll-out kernels) at the reference's training
shape: FLUX, bs 4 per GPU, latent 16x128x128 (S = 4608) — BASELINE.json configs[3] (forward part only this round)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200 import _lib  # noqa: E402
from arcflow_b200.config import flux_dev  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel, FluxTeacherEngine  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict, make_flux_teacher_extras  # noqa: E402
from arcflow_b200.train import ArcFlowDistillStep, draw_rollout_randoms  # noqa: E402

dev = torch.device("cuda", 0)
cfg = flux_dev()
sd = make_flux_state_dict(cfg, 1234, dev)
student = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
teacher = FluxTeacherEngine(student, make_flux_teacher_extras(cfg, 99, dev))
B = 4
x, txt, pooled = make_flux_inputs(cfg, B, 1024, 1024, 512, 42, dev)
step = ArcFlowDistillStep(student, teacher)
g = torch.Generator().manual_seed(0)
rands = [draw_rollout_randoms(B, 4, 16, g) for _ in range(2)]
for _ in range(2):
    step.forward(txt, pooled, (64, 64), x, rands, iteration=500)
torch.cuda.synchronize()
n0 = _lib.load().afb_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    loss, lv, _ = step.forward(txt, pooled, (64, 64), x, rands, iteration=500)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
flops = (2 * 78.77e12 + 8 * 74.36e12) * B
print(json.dumps(dict(what="train-step forward (no backward yet)", batch=B, ms=ms, loss=loss, tflops=flops / (ms * 1e9),
                      launches=(_lib.load().afb_launch_count() - n0) // 3, log_vars=lv)))
