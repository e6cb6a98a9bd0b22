"""Short driver for ncu: full-size ArcFlow-FLUX, one warm-up + one measured 2-NFE denoise step
(batch 8, 1024x1024). Depth can be reduced for `--set full` captures (kernels are identical per block)."""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200.config import ArcFluxConfig  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--double", type=int, default=19)
ap.add_argument("--single", type=int, default=38)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--px", type=int, default=1024)
ap.add_argument("--nfe", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--fuse-lora", action="store_true", help="merge the adapter into the base weights first (fuse_lora)")
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
cfg = ArcFluxConfig(num_layers=a.double, num_single_layers=a.single)
dev = torch.device("cuda", 0)
sd = make_flux_state_dict(cfg, seed=1234, device=dev)
model = ArcFluxEngineModel(sd, cfg, device=dev, consume_state_dict=True)
del sd
if a.fuse_lora:
    model.fuse_lora()
x, txt, pooled = make_flux_inputs(cfg, a.batch, a.px, a.px, device=dev)
grid = (a.px // 16, a.px // 16)
for _ in range(a.warmup):
    model.denoise(x, txt, pooled, grid, num_inference_steps=a.nfe)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    model.denoise(x, txt, pooled, grid, num_inference_steps=a.nfe)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"step {ms:.2f} ms  {a.batch / (ms / 1e3):.3f} images/s  fuse_lora={a.fuse_lora}")
