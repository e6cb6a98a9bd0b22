"""One checkpointed student forward + one adapter backward on a FLUX-width model of reduced depth (2 double + 2 single
blocks, B=4, 1024^2) — small enough to run under `ncu --metrics gpu__time_duration.sum` for a per-kernel split of the
backward. Prints CUDA-event times when run plainly."""
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200.config import flux_dev  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
cfg = flux_dev()
cfg.num_layers, cfg.num_single_layers = 2, 2
sd = make_flux_state_dict(cfg, 1234, dev)
student = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
x, txt, pooled = make_flux_inputs(cfg, B, 1024, 1024, 512, 42, dev)
grads = {n: torch.zeros(s, dtype=torch.float32, device=dev) for n, s in
         {**student.trunk_lora_shapes(), **student.embed_lora_shapes()}.items()}
dy = torch.randn(B, 4096, cfg.inner_dim, device=dev).mul(1e-3).bfloat16()
d_mod = torch.zeros(B, student.mod_total, dtype=torch.float32, device=dev)


def run(with_mod):
    student.forward_heads(x, txt, pooled, 0.8, 3.5, (64, 64), train=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    student.backward_trunk(dy, grads, d_mod if with_mod else None)
    if with_mod:
        student.backward_embed(d_mod, grads)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


run(True)
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
student.forward_heads(x, txt, pooled, 0.8, 3.5, (64, 64))
f1.record()
torch.cuda.synchronize()
print(json.dumps(dict(batch=B, blocks="2 double + 2 single", forward_ms=f0.elapsed_time(f1), backward_ms=run(False),
                      backward_with_mod_grads_ms=run(True))))
