"""Times the native train iteration at the reference's training shape: FLUX, bs 4 per GPU, latent 16x128x128
(S_img = 4096, S_txt = 512) — BASELINE.json configs[3]: 2 student + 8 teacher forwards, roll-out kernels, the full
adapter backward (per-block recompute), grad clip + AdamW + EMA and the bf16 write-back.
Usage: python tools/train_step_time.py [batch] [--profile]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200 import _lib  # noqa: E402
from arcflow_b200.config import flux_dev  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel, FluxTeacherEngine  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict, make_flux_teacher_extras  # noqa: E402
from arcflow_b200.train import ArcFlowTrainer, draw_rollout_randoms  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(args[0]) if args else 4
profile = "--profile" in sys.argv
dev = torch.device("cuda", 0)
cfg = flux_dev()
sd = make_flux_state_dict(cfg, 1234, dev)
student = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
del sd
teacher = FluxTeacherEngine(student, make_flux_teacher_extras(cfg, 99, dev))
x, txt, pooled = make_flux_inputs(cfg, B, 1024, 1024, 512, 42, dev)
trainer = ArcFlowTrainer(student, teacher)
step = trainer.distill
g = torch.Generator().manual_seed(0)
rands = [draw_rollout_randoms(B, 4, 16, g) for _ in range(2)]
lib = _lib.load()


def timed(fn, n):
    torch.cuda.synchronize()
    n0 = lib.afb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (lib.afb_launch_count() - n0) // n, out


trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500)   # warm-up (also sizes every workspace)
fwd_ms, fwd_launches, (loss_f, _, _) = timed(lambda: step.forward(txt, pooled, (64, 64), x, rands, iteration=500), 2)
ms, launches, (loss, lv) = timed(lambda: trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500), 2)
fwd_flops = (2 * 78.77e12 + 8 * 74.36e12) * B
out = dict(what="train iteration (fwd + bwd + optimizer)", batch=B, ms=ms, forward_only_ms=fwd_ms, backward_optim_ms=ms - fwd_ms,
           samples_per_s=B / (ms / 1e3), loss=loss, fwd_tflops=fwd_flops / (fwd_ms * 1e9), launches=launches,
           fwd_launches=fwd_launches, mem_gb=torch.cuda.max_memory_allocated() / 2**30,
           log_vars={k: (float(v) if isinstance(v, (int, float)) else v) for k, v in lv.items()})
if profile:
    student.set_profiling(True)
    trainer.train_step(txt, pooled, (64, 64), x, rands, iteration=500)
    out["student_profile"] = student.read_profile()
    student.set_profiling(False)
print(json.dumps(out))
