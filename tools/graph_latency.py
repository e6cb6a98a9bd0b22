"""Latency of one 2-NFE FLUX generation, eager launch sequence vs the captured CUDA graph, at small shapes where the
loop is launch-bound: python tools/graph_latency.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200.config import flux_dev  # noqa: E402
from arcflow_b200.model import ArcFluxEngineModel  # noqa: E402
from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
cfg = flux_dev()
model = ArcFluxEngineModel(make_flux_state_dict(cfg, 1234, dev), cfg, dev, consume_state_dict=True)
for B, px in [(1, 256), (1, 512), (1, 1024), (4, 512)]:
    x, txt, pooled = make_flux_inputs(cfg, B, px, px, 512, 42, dev)
    grid = (px // 16, px // 16)
    res = {}
    for mode in (False, True):
        for _ in range(3):
            model.denoise(x, txt, pooled, grid, cuda_graph=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            model.denoise(x, txt, pooled, grid, cuda_graph=mode)
        e1.record()
        torch.cuda.synchronize()
        res["graph_ms" if mode else "eager_ms"] = e0.elapsed_time(e1) / n
    print(json.dumps(dict(batch=B, px=px, nfe=2, **res, speedup=res["eager_ms"] / res["graph_ms"])))
